// Device math for the spectral path-tracing hot path: shading-space vectors, sampling, Fresnel, GGX
// microfacets, tangent frames.  fp32 throughout; the library is built with -fmad=false -prec-div=true
// -prec-sqrt=true -ftz=true so that +,-,*,/,sqrt round exactly like the scalar host code of the reference
// (FTZ/DAZ, reference src/base/Platform.h:20-34); fused multiply-adds appear only where the reference
// itself calls std::fma.  Reference file:line cited per function.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace prb {
#define PRB_DEV __device__ __forceinline__
// out-of-line: the microfacet / Fresnel kernels are called from many places (6 materials x eval/pdf/sample x 4
// wavelengths); inlining every copy made k_shade ~1 MB of SASS and instruction-fetch bound (ncu: stall_no_instruction
// 63 warps per issue).  One copy each keeps the hot code inside the instruction cache.
#define PRB_DEV_NI __device__ __noinline__
// Code-size diet of k_shade (round 2): ncu shows the shading kernel INSTRUCTION-FETCH bound (sm__icc_request_hit_rate 76 %,
// gcc instruction requests at 55 % of peak, stall_no_instruction 2.6 warps per issue; more resident warps change nothing) --
// 219 KB of SASS for the all-Lambert instantiation, 22 % of it seven inlined copies of the double-precision sincos.  Level 1
// takes the correctly rounded transcendental wrappers out of line, level 2 also the medium-sized helpers that are inlined
// at many call sites (normalized, tableLookup, evalLeafNode, fragmentXYZ); results are bit-identical by construction.
#ifndef PRB_SMALL_CODE
#define PRB_SMALL_CODE 0
#endif
#if PRB_SMALL_CODE >= 1
#define PRB_DEV_TRANS __device__ __noinline__
#else
#define PRB_DEV_TRANS __device__ __forceinline__
#endif
#if PRB_SMALL_CODE >= 2
#define PRB_DEV_MED __device__ __noinline__
#else
#define PRB_DEV_MED __device__ __forceinline__
#endif

constexpr float PR_EPSILON	= 1.1920928955078125e-07f; // std::numeric_limits<float>::epsilon()
constexpr float PR_PI		= 3.14159265358979323846f;
constexpr float PR_INV_PI	= 0.31830988618379067154f;
constexpr float PR_INV_2_PI = 0.15915494309189533577f;
#define PRB_INF CUDART_INF_F

struct V3 {
	float x, y, z;
};
PRB_DEV V3 mk(float x, float y, float z) { return V3{ x, y, z }; }
PRB_DEV V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
PRB_DEV V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
PRB_DEV V3 operator-(V3 a) { return { -a.x, -a.y, -a.z }; }
PRB_DEV V3 operator*(V3 a, float f) { return { a.x * f, a.y * f, a.z * f }; }
PRB_DEV V3 operator*(float f, V3 a) { return { a.x * f, a.y * f, a.z * f }; }
PRB_DEV V3 operator/(V3 a, float f) { return { a.x / f, a.y / f, a.z / f }; }
PRB_DEV float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
PRB_DEV V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
PRB_DEV float norm2(V3 a) { return dot(a, a); }
// x / s for s > 0 (not NaN), bit-identical to the IEEE division.  The division sequence (reciprocal + Newton steps + FCHK)
// leaves through a ~40-instruction out-of-line slow path whenever its numerator is ZERO, and the normals and tangents of
// axis-aligned faces (every wall of a Cornell box) have two zero components: ncu counted 14 such calls per warp in k_shade,
// 11 % of its instructions (profiles/r02_kshade_experiments.txt, item 7).  +-0 / s = +-0 for s > 0, so a zero numerator
// skips the division; `asm volatile` keeps the compiler from speculating the division above the branch.
PRB_DEV float divPositive(float x, float s)
{
	if (x != 0.0f)
		asm volatile("div.rn.ftz.f32 %0, %0, %1;" : "+f"(x) : "f"(s));
	return x;
}
PRB_DEV_MED V3 normalized(V3 a)
{
	const float z = norm2(a);
	if (!(z > 0))
		return a;
	const float s = sqrtf(z);
	return { divPositive(a.x, s), divPositive(a.y, s), divPositive(a.z, s) };
}
PRB_DEV bool isZero(V3 a, float prec) { return fabsf(a.x) <= prec && fabsf(a.y) <= prec && fabsf(a.z) <= prec; }
PRB_DEV V3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }

// SpectralBlob (reference src/core/spectral/SpectralBlob.h:7-20): four wavelengths kept in registers
struct Blob {
	float v[4];
	PRB_DEV float& operator[](int i) { return v[i]; }
	PRB_DEV float operator[](int i) const { return v[i]; }
};
PRB_DEV Blob blob(float f) { return Blob{ { f, f, f, f } }; }
PRB_DEV Blob blob4(float4 f) { return Blob{ { f.x, f.y, f.z, f.w } }; }
PRB_DEV float4 tof4(Blob b) { return make_float4(b[0], b[1], b[2], b[3]); }
PRB_DEV Blob operator*(Blob a, Blob b) { return { { a[0] * b[0], a[1] * b[1], a[2] * b[2], a[3] * b[3] } }; }
PRB_DEV Blob operator*(Blob a, float f) { return { { a[0] * f, a[1] * f, a[2] * f, a[3] * f } }; }
PRB_DEV Blob operator/(Blob a, Blob b) { return { { a[0] / b[0], a[1] / b[1], a[2] / b[2], a[3] / b[3] } }; }
PRB_DEV Blob operator/(Blob a, float f) { return { { a[0] / f, a[1] / f, a[2] / f, a[3] / f } }; }
PRB_DEV Blob operator+(Blob a, Blob b) { return { { a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3] } }; }
PRB_DEV float bsum(Blob a) { return ((a[0] + a[1]) + a[2]) + a[3]; }
PRB_DEV Blob heroOnly() { return Blob{ { 1.0f, 0.0f, 0.0f, 0.0f } }; }
PRB_DEV bool allLE(Blob a, float e) { return a[0] <= e && a[1] <= e && a[2] <= e && a[3] <= e; }
PRB_DEV bool blobIsZero(Blob a, float e) { return fabsf(a[0]) <= e && fabsf(a[1]) <= e && fabsf(a[2]) <= e && fabsf(a[3]) <= e; }

// reference src/base/config/MathGlue.inl:8-24
PRB_DEV float diffProd(float a, float b, float c, float d)
{
	const float cd	= c * d;
	const float err = fmaf(-c, d, cd);
	const float dop = fmaf(a, b, -cd);
	return dop + err;
}
PRB_DEV float sumProd(float a, float b, float c, float d) { return fmaf(a, b, c * d); }

// ---------------------------------------------------------------- ShadingVector (src/base/math/ShadingVector.h)
PRB_DEV bool signbitf(float f) { return (__float_as_uint(f) >> 31) != 0; }
PRB_DEV float cosTheta(V3 v) { return v.z; }
PRB_DEV float cos2Theta(V3 v) { return v.z * v.z; }
PRB_DEV float absCosTheta(V3 v) { return fabsf(v.z); }
PRB_DEV float sin2Theta(V3 v) { return fmaxf(0.0f, 1 - cos2Theta(v)); }
PRB_DEV float tan2Theta(V3 v) { return absCosTheta(v) <= PR_EPSILON ? 0 : sin2Theta(v) / cos2Theta(v); }
PRB_DEV_NI float cos2Phi(V3 v)
{
	const float s = sin2Theta(v);
	return s <= PR_EPSILON ? 0 : fminf(1.0f, v.x * v.x / s);
}
PRB_DEV_NI float sin2Phi(V3 v)
{
	const float s = sin2Theta(v);
	return s <= PR_EPSILON ? 0 : fminf(1.0f, v.y * v.y / s);
}
PRB_DEV bool sameHemisphere(V3 a, V3 b) { return signbitf(a.z) == signbitf(b.z); }
PRB_DEV bool isPositiveHemisphere(V3 a) { return !signbitf(a.z); }
PRB_DEV V3 makeSameHemisphere(V3 self, V3 other) { return sameHemisphere(self, other) ? other : -other; }
PRB_DEV V3 makePositiveHemisphere(V3 a) { return isPositiveHemisphere(a) ? a : -a; }

// IEEE division that survives the optimiser: under -ftz=true NVVM rewrites `x / constant` into `x * (1 / constant)`
// (1 ulp off for constants that are not powers of two) even with -prec-div=true; the intrinsic is left alone.  Used
// wherever the divisor is, or can become after inlining, a compile-time constant (checked by tools/check_ptx_div.sh).
PRB_DEV float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// ---------------------------------------------------------------- transcendental functions
// The reference calls std::sin/cos/tan/atan/atan2/acos on float, i.e. whatever libm the build links (glibc's are within
// 1 ulp but not correctly rounded: 1.3 % of sinf and 15 % of atan2f results differ from the correctly rounded value),
// so the last bit is libm-version dependent.  Device and oracle both use the CORRECTLY ROUNDED fp32 result instead
// (evaluate in fp64, round once): libm-independent, and the two sides agree bit for bit, which keeps long specular
// chains from diverging chaotically.
PRB_DEV_TRANS void cr_sincos(float x, float* s, float* c)
{
	double ds, dc;
	sincos((double)x, &ds, &dc);
	*s = (float)ds;
	*c = (float)dc;
}
PRB_DEV_TRANS float cr_sin(float x) { return (float)sin((double)x); }
PRB_DEV_TRANS float cr_cos(float x) { return (float)cos((double)x); }
PRB_DEV_TRANS float cr_tan(float x) { return (float)tan((double)x); }
PRB_DEV_TRANS float cr_atanh(float x) { return (float)atanh((double)x); }
PRB_DEV_TRANS float cr_cosh(float x) { return (float)cosh((double)x); }
PRB_DEV_TRANS float cr_atan(float x) { return (float)atan((double)x); }
PRB_DEV_TRANS float cr_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
PRB_DEV_TRANS float cr_acos(float x) { return (float)acos((double)x); }

// ---------------------------------------------------------------- Sampling (src/base/math/Sampling.h:38-57)
PRB_DEV V3 cos_hemi(float u1, float u2)
{
	const float cosT = sqrtf(u1);
	const float sinT = sqrtf(1 - u1);
	const float phi	 = 2 * PR_PI * u2;
	float sinPhi, cosPhi;
	cr_sincos(phi, &sinPhi, &cosPhi);
	return mk(sinT * cosPhi, sinT * sinPhi, cosT);
}
PRB_DEV float cos_hemi_pdf(float NdotL) { return NdotL * PR_INV_PI; }

// ---------------------------------------------------------------- Scattering (src/base/math/Scattering.h:49-183)
PRB_DEV_NI float refraction_angle(float cosI, float eta)
{
	if (signbitf(cosI)) {
		cosI = -cosI;
		eta	 = 1 / eta;
	}
	const float k = 1 - (eta * eta) * (1 - cosI * cosI);
	return k < 0 ? -1.0f : sqrtf(k);
}
PRB_DEV V3 reflectZ(V3 V) { return mk(-V.x, -V.y, V.z); }
PRB_DEV V3 reflectN(V3 V, V3 N) { return (2 * dot(N, V)) * N - V; }
PRB_DEV V3 refractZ(float eta, V3 wIn)
{
	const bool neg = signbitf(wIn.z);
	if (neg) {
		eta = 1 / eta;
		wIn = -wIn;
	}
	const float cosT = refraction_angle(wIn.z, eta);
	V3 r			 = cosT < 0.0f ? reflectZ(wIn) : normalized(mk(-wIn.x * eta, -wIn.y * eta, -cosT));
	return neg ? -r : r;
}
PRB_DEV_NI V3 refractN(float eta, V3 wIn, V3 N, bool& total)
{
	float cosI	   = dot(wIn, N);
	const bool neg = signbitf(cosI);
	if (neg) { // -refract(1/eta, -wIn, N, total)
		eta	 = 1 / eta;
		wIn	 = -wIn;
		cosI = dot(wIn, N);
	}
	const float cosT = refraction_angle(cosI, eta);
	total			 = cosT < 0.0f;
	V3 r			 = total ? reflectN(wIn, N) : normalized((-wIn) * eta + (eta * cosI - cosT) * N);
	return neg ? -r : r;
}
PRB_DEV V3 halfway_reflection(V3 wIn, V3 wOut) { return normalized(wIn + wOut); }
PRB_DEV V3 halfway_refractive(float n_in, V3 wIn, float n_out, V3 wOut) { return -normalized(n_in * wIn + n_out * wOut); }
PRB_DEV float reflective_jacobian(float cosO)
{
	const float denom = 4 * fabsf(cosO);
	return denom <= PR_EPSILON ? 0.0f : 1 / denom;
}
PRB_DEV_NI float refractive_jacobian(float eta, float cosI, float cosO)
{
	const float denom  = eta * cosI + cosO;
	const float denom2 = denom * denom;
	return denom2 <= PR_EPSILON ? 0.0f : fabsf(cosO) / denom2;
}

// ---------------------------------------------------------------- Fresnel (src/base/math/Fresnel.h:9-77)
PRB_DEV_NI float fresnel_dielectric(float cosI, float n_in, float n_out)
{
	if (signbitf(cosI)) { // negative hemisphere: dielectric(-cosI, n_out, n_in)
		cosI		  = -cosI;
		const float t = n_in;
		n_in		  = n_out;
		n_out		  = t;
	}
	const float cosT = refraction_angle(cosI, n_in / n_out);
	if (cosT < 0)
		return 1;
	const float perp = diffProd(n_in, cosI, n_out, cosT) / sumProd(n_in, cosI, n_out, cosT);
	const float para = diffProd(n_out, cosI, n_in, cosT) / sumProd(n_out, cosI, n_in, cosT);
	return fminf(fmaxf(sumProd(para, para, perp, perp) / 2.0f, 0.0f), 1.0f);
}
PRB_DEV_NI float fresnel_conductor(float cosI, float n_in, float n_out, float k)
{
	if (cosI < 0)
		cosI = -cosI;
	const float eta	   = n_out / n_in;
	const float kappa  = k / n_in;
	const float cosI2  = cosI * cosI;
	const float sinI2  = 1 - cosI2;
	const float eta2   = eta * eta;
	const float kappa2 = kappa * kappa;
	const float t0	   = eta2 - kappa2 - sinI2;
	const float ap	   = sqrtf(sumProd(t0, t0, 4 * eta2, kappa2));
	const float t1	   = ap + cosI2;
	const float a	   = sqrtf((ap + t0) / 2);
	const float t2	   = 2 * cosI * a;
	const float perp2  = (t1 - t2) / (t1 + t2);
	const float t3	   = sumProd(cosI2, ap, sinI2, sinI2);
	const float t4	   = t2 * sinI2;
	const float para2  = perp2 * (t3 - t4) / (t3 + t4);
	const float R	   = (para2 + perp2) / 2;
	return fminf(fmaxf(R, 0.0f), 1.0f);
}
PRB_DEV float schlick_term(float d)
{
	const float t = 1 - d;
	return (t * t) * (t * t) * t;
}
PRB_DEV float schlick(float d, float f0) { return f0 + (1 - f0) * schlick_term(d); }

// ---------------------------------------------------------------- Microfacet (src/base/math/Microfacet.h)
PRB_DEV_NI float g_1_smith_opt(float NdotK, float roughness)
{
	const float a	  = roughness * roughness;
	const float b	  = NdotK * NdotK;
	const float denom = NdotK + sqrtf(a + b - a * b);
	return (denom <= PR_EPSILON) ? 0.0f : 1.0f / denom;
}
PRB_DEV_NI float g_1_smith1(V3 K, float roughness)
{
	const float a	  = roughness * roughness;
	const float b	  = tan2Theta(K);
	const float denom = 1 + sqrtf(1 + a * b);
	return (denom <= PR_EPSILON) ? 0.0f : 2.0f / denom;
}
PRB_DEV_NI float g_1_smith2(V3 K, float rx, float ry)
{
	const float ax2	  = cos2Phi(K) * rx * rx;
	const float ay2	  = sin2Phi(K) * ry * ry;
	const float b	  = tan2Theta(K);
	const float denom = 1 + sqrtf(1 + (ax2 + ay2) * b);
	return (denom <= PR_EPSILON) ? 0.0f : 2.0f / denom;
}
PRB_DEV_NI float g_1_smith_lambda1(V3 K, float roughness)
{
	const float a = roughness * roughness;
	const float b = tan2Theta(K);
	return (sqrtf(1 + a * b) - 1) / 2;
}
PRB_DEV_NI float g_1_smith_lambda2(V3 K, float rx, float ry)
{
	const float ax2 = cos2Phi(K) * rx * rx;
	const float ay2 = sin2Phi(K) * ry * ry;
	const float b	= tan2Theta(K);
	return (sqrtf(1 + (ax2 + ay2) * b) - 1) / 2;
}
PRB_DEV_NI float ndf_ggx1(V3 H, float roughness)
{
	const float sin2 = sin2Theta(H);
	const float cos2 = cos2Theta(H);
	if (cos2 <= PR_EPSILON)
		return 0.0f;
	const float tan2   = sin2 / cos2;
	const float cos4   = cos2 * cos2;
	const float alpha2 = roughness * roughness;
	if (alpha2 <= PR_EPSILON)
		return 0.0f;
	const float e	  = tan2 / alpha2;
	const float denom = alpha2 * cos4 * (1 + e) * (1 + e);
	return (denom <= PR_EPSILON) ? 0.0f : PR_INV_PI / denom;
}
PRB_DEV_NI float ndf_ggx2(V3 H, float rx, float ry)
{
	const float sin2 = sin2Theta(H);
	const float cos2 = cos2Theta(H);
	if (cos2 <= PR_EPSILON)
		return 0.0f;
	const float tan2	= sin2 / cos2;
	const float cos4	= cos2 * cos2;
	const float alphaX2 = rx * rx;
	const float alphaY2 = ry * ry;
	if (alphaX2 <= PR_EPSILON || alphaY2 <= PR_EPSILON)
		return 0.0f;
	const float t	  = sin2Phi(H) / alphaX2 + cos2Phi(H) / alphaY2;
	const float e	  = tan2 * t;
	const float denom = rx * ry * cos4 * (1 + e) * (1 + e);
	return (denom <= PR_EPSILON) ? 0.0f : PR_INV_PI / denom;
}
PRB_DEV V3 spherical_cartesian(float thSin, float thCos, float phSin, float phCos) { return mk(thSin * phCos, thSin * phSin, thCos); }
PRB_DEV_NI V3 sample_ndf_ggx1(float u0, float u1, float roughness)
{
	const float alpha2 = roughness * roughness;
	const float t2	   = alpha2 * u1 / (1 - u1);
	const float cosT   = alpha2 <= PR_EPSILON ? 1.0f : fmaxf(0.001f, 1.0f / sqrtf(1 + t2));
	const float sinT   = sqrtf(1 - cosT * cosT);
	float sinPhi, cosPhi;
	cr_sincos(2 * PR_PI * u0, &sinPhi, &cosPhi);
	return spherical_cartesian(sinT, cosT, sinPhi, cosPhi);
}
PRB_DEV_NI V3 sample_ndf_ggx2(float u0, float u1, float rx, float ry)
{
	const float phi = cr_atan(ry / rx * cr_tan(PR_PI + 2 * PR_PI * u0)) + PR_PI * floorf(2 * u0 + 0.5f);
	float sinPhi, cosPhi;
	cr_sincos(phi, &sinPhi, &cosPhi);
	const float f1	   = cosPhi / rx;
	const float f2	   = sinPhi / ry;
	const float alpha2 = 1 / (f1 * f1 + f2 * f2);
	const float t2	   = alpha2 * u1 / (1 - u1);
	const float cosT   = fmaxf(0.001f, 1.0f / sqrtf(1 + t2));
	const float sinT   = sqrtf(1 - cosT * cosT);
	return spherical_cartesian(sinT, cosT, sinPhi, cosPhi);
}
PRB_DEV_NI float pdf_ggx_vndf(V3 V, V3 H, float rx, float ry)
{
	return absCosTheta(V) <= PR_EPSILON ? 0.0f : g_1_smith2(V, rx, ry) * fabsf(dot(V, H)) * ndf_ggx2(H, rx, ry) / absCosTheta(V);
}
PRB_DEV_NI V3 sample_vndf_ggx(float u0, float u1, V3 nV, float rx, float ry)
{ // Heitz 2018, Microfacet.h:261-330 (#if 1 branch)
	const V3 Vh		  = normalized(mk(rx * nV.x, ry * nV.y, nV.z));
	const float lensq = sumProd(Vh.x, Vh.x, Vh.y, Vh.y);
	const V3 T1		  = lensq > PR_EPSILON ? mk(-Vh.y, Vh.x, 0) / sqrtf(lensq) : mk(1, 0, 0);
	const V3 T2		  = cross(Vh, T1);
	const float r	  = sqrtf(u0);
	const float phi	  = 2.0f * PR_PI * u1;
	float sp, cp;
	cr_sincos(phi, &sp, &cp);
	const float t1 = r * cp;
	float t2	   = r * sp;
	const float s  = 0.5f * (1.0f + Vh.z);
	t2			   = (1.0f - s) * sqrtf(1.0f - t1 * t1) + s * t2;
	const V3 Nh	   = t1 * T1 + t2 * T2 + sqrtf(fmaxf(0.0f, 1.0f + diffProd(-t1, t1, t2, t2))) * Vh;
	return normalized(mk(rx * Nh.x, ry * Nh.y, fmaxf(0.0f, Nh.z)));
}

// ---------------------------------------------------------------- RoughDistribution / MicrofacetReflection / -Transmission
struct RoughDistribution { // src/base/math/RoughDistribution.h
	float M1, M2;
	bool aniso, vndf;
	PRB_DEV bool isDelta() const { return M1 <= 1e-3f || M2 <= 1e-3f; }
	PRB_DEV_NI float G(V3 H, V3 V, V3 L) const
	{
		const bool chi_v = cosTheta(V) * dot(H, V) > PR_EPSILON;
		const bool chi_l = cosTheta(L) * dot(H, L) > PR_EPSILON;
		if (!chi_v || !chi_l)
			return 0.0f;
		if (!vndf)
			return aniso ? g_1_smith2(V, M1, M2) * g_1_smith2(L, M1, M2) : g_1_smith1(V, M1) * g_1_smith1(L, M1);
		const float denom = aniso ? 1 + g_1_smith_lambda2(V, M1, M2) + g_1_smith_lambda2(L, M1, M2) : 1 + g_1_smith_lambda1(V, M1) + g_1_smith_lambda1(L, M1);
		return denom <= PR_EPSILON ? 0.0f : 1 / denom;
	}
	PRB_DEV float D(V3 H) const { return aniso ? ndf_ggx2(H, M1, M2) : ndf_ggx1(H, M1); }
	PRB_DEV float Norm(V3 H, V3 V, V3 L) const
	{
		const float denom = absCosTheta(V);
		if (denom <= PR_EPSILON)
			return 0;
		return fabsf(dot(H, L)) / denom;
	}
	PRB_DEV float DGNorm(V3 H, V3 V, V3 L) const { return D(H) * G(H, V, L) * Norm(H, V, L); }
	PRB_DEV_NI float pdf(V3 H, V3 V) const
	{
		if (isDelta())
			return 1.0f;
		if (vndf)
			return pdf_ggx_vndf(makePositiveHemisphere(V), makePositiveHemisphere(H), M1, M2);
		return (aniso ? ndf_ggx2(H, M1, M2) : ndf_ggx1(H, M1)) * absCosTheta(H);
	}
	PRB_DEV_NI V3 sample(float r0, float r1, V3 V) const
	{
		if (isDelta())
			return mk(0, 0, 1);
		if (vndf)
			return sample_vndf_ggx(r0, r1, makePositiveHemisphere(V), M1, M2);
		return aniso ? sample_ndf_ggx2(r0, r1, M1, M2) : sample_ndf_ggx1(r0, r1, M1);
	}
};
struct MicrofacetReflection { // src/base/math/MicrofacetReflection.h
	RoughDistribution D;
	PRB_DEV bool isDelta() const { return D.isDelta(); }
	PRB_DEV_NI float evalDielectric(V3 wIn, V3 wOut, float n_in, float n_out) const
	{
		if (!sameHemisphere(wIn, wOut))
			return 0.0f;
		V3 H = halfway_reflection(wIn, wOut);
		if (!isPositiveHemisphere(H))
			H = -H;
		const float cosI = dot(H, wIn);
		const float F	 = fresnel_dielectric(cosI, n_in, n_out);
		if (isDelta())
			return F;
		return F * D.DGNorm(H, wIn, wOut) * reflective_jacobian(cosI);
	}
	PRB_DEV_NI float evalConductor(V3 wIn, V3 wOut, float ior, float kappa) const
	{
		if (!sameHemisphere(wIn, wOut))
			return 0.0f;
		V3 H = halfway_reflection(wIn, wOut);
		if (!isPositiveHemisphere(H))
			H = -H;
		const float cosI = dot(H, wIn);
		const float F	 = fresnel_conductor(cosI, 1, ior, kappa);
		if (isDelta())
			return F;
		return F * D.DGNorm(H, wIn, wOut) * reflective_jacobian(cosI);
	}
	PRB_DEV_NI float eval(V3 wIn, V3 wOut) const
	{
		if (!sameHemisphere(wIn, wOut))
			return 0.0f;
		const V3 H = halfway_reflection(wIn, wOut);
		if (isDelta())
			return 1.0f;
		return D.DGNorm(H, wIn, wOut) * reflective_jacobian(dot(H, wIn));
	}
	PRB_DEV_NI float pdf(V3 wIn, V3 wOut) const
	{
		if (!sameHemisphere(wIn, wOut))
			return 0.0f;
		const V3 H = halfway_reflection(wIn, wOut);
		if (isDelta())
			return 1.0f;
		return reflective_jacobian(dot(H, wIn)) * D.pdf(H, wIn);
	}
	PRB_DEV_NI V3 sample(float r0, float r1, V3 wIn) const
	{
		const V3 H = D.sample(r0, r1, wIn);
		if (isZero(H, PR_EPSILON))
			return mk(0, 0, 0);
		const V3 wOut = reflectN(wIn, H);
		return sameHemisphere(wIn, wOut) ? wOut : mk(0, 0, 0);
	}
};
struct MicrofacetTransmission { // src/base/math/MicrofacetTransmission.h
	RoughDistribution D;
	float InnerIOR, OuterIOR;
	PRB_DEV bool isDelta() const { return D.isDelta(); }
	PRB_DEV_NI bool setup(V3 wIn, V3 wOut, V3& H, float& cosI, float& cosO, float& eta) const
	{
		if (sameHemisphere(wIn, wOut))
			return false;
		const float in_ior	= isPositiveHemisphere(wIn) ? InnerIOR : OuterIOR;
		const float out_ior = isPositiveHemisphere(wIn) ? OuterIOR : InnerIOR;
		H					= halfway_refractive(in_ior, wIn, out_ior, wOut);
		if (!isPositiveHemisphere(H))
			H = -H;
		cosI = dot(H, wIn);
		cosO = dot(H, wOut);
		if (cosI * cosO >= -PR_EPSILON)
			return false;
		eta = in_ior / out_ior;
		return true;
	}
	PRB_DEV_NI float evalDielectric(V3 wIn, V3 wOut, bool isLightPath) const
	{
		V3 H;
		float cosI, cosO, eta;
		if (!setup(wIn, wOut, H, cosI, cosO, eta))
			return 0.0f;
		const float F = fresnel_dielectric(cosI, InnerIOR, OuterIOR);
		if (isDelta())
			return 1 - F;
		const float jacobian = refractive_jacobian(eta, cosI, cosO);
		const float spread	 = isLightPath ? 1 / (eta * eta) : 1.0f;
		return (1 - F) * D.DGNorm(H, wIn, wOut) * jacobian * spread;
	}
	PRB_DEV_NI float pdf(V3 wIn, V3 wOut) const
	{
		V3 H;
		float cosI, cosO, eta;
		if (!setup(wIn, wOut, H, cosI, cosO, eta))
			return 0.0f;
		if (isDelta())
			return 1.0f;
		return D.pdf(H, wIn) * refractive_jacobian(eta, cosI, cosO);
	}
	PRB_DEV_NI V3 sample(float r0, float r1, V3 wIn) const
	{
		const V3 H = D.sample(r0, r1, wIn);
		if (isZero(H, PR_EPSILON))
			return mk(0, 0, 0);
		bool total;
		const V3 L = refractN(InnerIOR / OuterIOR, wIn, H, total);
		return (total == sameHemisphere(wIn, L)) ? L : mk(0, 0, 0);
	}
};

// ---------------------------------------------------------------- Tangent.h:9-56, Transform.h:9-32, Spherical.h
PRB_DEV V3 fromTangentSpace(V3 N, V3 Nx, V3 Ny, V3 V) { return normalized((N * V.z + Ny * V.y) + Nx * V.x); }
PRB_DEV V3 toTangentSpace(V3 N, V3 Nx, V3 Ny, V3 V) { return normalized(mk(dot(Nx, V), dot(Ny, V), dot(N, V))); }
PRB_DEV void frame_duff(V3 N, V3& Nx, V3& Ny)
{
	const float sign = copysignf(1.0f, N.z);
	const float a	 = -1.0f / (sign + N.z);
	const float b	 = N.x * N.y * a;
	Nx				 = mk(1.0f + sign * N.x * N.x * a, sign * b, -sign * N.x);
	Ny				 = mk(b, sign + N.y * N.y * a, -N.y);
}
PRB_DEV void tangent_frame(V3 N, V3& Nx, V3& Ny)
{
	frame_duff(N, Nx, Ny);
	Nx = normalized(Nx);
	Ny = normalized(Ny);
}
PRB_DEV float nextFloatUp(float v)
{ // src/base/config/Types.inl:140-152
	if (isinf(v) && v > 0.0f)
		return v;
	if (v == -0.0f)
		v = 0.0f;
	uint32_t ui = __float_as_uint(v);
	if (v >= 0)
		++ui;
	else
		--ui;
	return __uint_as_float(ui);
}
PRB_DEV float nextFloatDown(float v)
{
	if (isinf(v) && v < 0.0f)
		return v;
	if (v == 0.0f)
		v = -0.0f;
	uint32_t ui = __float_as_uint(v);
	if (v > 0)
		--ui;
	else
		++ui;
	return __uint_as_float(ui);
}
PRB_DEV float adjustUlp(float p, float off) { return off > 0 ? nextFloatUp(p) : (off < 0 ? nextFloatDown(p) : p); }
PRB_DEV V3 safePosition(V3 pos, V3 dir, V3 N)
{
	const float d = ((fabsf(N.x) * 0.0001f + fabsf(N.y) * 0.0001f) + fabsf(N.z) * 0.0001f);
	V3 offset	  = d * N;
	if (dot(dir, N) < 0)
		offset = -offset;
	const V3 p = pos + offset;
	return mk(adjustUlp(p.x, offset.x), adjustUlp(p.y, offset.y), adjustUlp(p.z, offset.z));
}
PRB_DEV void uv_from_normal(V3 N, float& u, float& v)
{ // Spherical::uv_from_normal / from_direction, src/base/math/Spherical.h:9-31
	const float x = (N.x == 0 && N.y == 0) ? 1e-5f : N.x;
	float phi	  = cr_atan2(N.y, x);
	phi			  = phi < 0 ? phi + 2 * PR_PI : phi;
	const float theta = cr_acos(N.z);
	const float tx = theta * PR_INV_PI, ty = phi * PR_INV_PI;
	u = ty / 2;
	v = tx;
}
PRB_DEV V3 cartesian_from_uv(float u, float v)
{
	const float theta = v * PR_PI, phi = u * 2 * PR_PI;
	float st, ct, sp, cp;
	cr_sincos(theta, &st, &ct);
	cr_sincos(phi, &sp, &cp);
	return spherical_cartesian(st, ct, sp, cp);
}

// ---------------------------------------------------------------- RNG (src/core/Random.h:26-179, pcg32_fast)
struct Rng {
	uint64_t s;
	PRB_DEV uint32_t get32()
	{
		const uint64_t old = s;
		s				   = old * 6364136223846793005ULL;
		const uint32_t rs  = (uint32_t)(old >> 61);
		const uint64_t x   = old ^ (old >> 22);
		return (uint32_t)(x >> (22 + rs));
	}
	PRB_DEV float getFloat() { return __uint_as_float((get32() >> 9) | 0x3F800000u) - 1.0f; }
	// Vector2f(getFloat(), getFloat()): first draw -> y, second -> x (GCC argument order)
	PRB_DEV void get2D(float& x, float& y)
	{
		y = getFloat();
		x = getFloat();
	}
};
} // namespace prb
