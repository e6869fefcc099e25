// ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the PearRay (reference @ 95f065a) spectral
// path-tracing hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use anything under oracle/; the product (pearray_b200/) never includes, links or calls it.
//
// This header: scalar fp32 math restated from the reference's base/math headers.  Each function cites the
// reference file:line it follows.  Compile with -ffp-contract=off: fused multiply-adds appear only where the
// reference calls std::fma explicitly.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {
constexpr float PR_EPSILON	= std::numeric_limits<float>::epsilon(); // src/base/config/Constants.inl
constexpr float PR_INF		= std::numeric_limits<float>::infinity();
constexpr float PR_PI		= 3.14159265358979323846f;
constexpr float PR_INV_PI	= 0.31830988618379067154f;
constexpr float PR_INV_2_PI = 0.15915494309189533577f;

struct V3 {
	float x, y, z;
};
inline V3 mk(float x, float y, float z) { return V3{ x, y, z }; }
inline V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 operator-(V3 a) { return { -a.x, -a.y, -a.z }; }
inline V3 operator*(V3 a, float f) { return { a.x * f, a.y * f, a.z * f }; }
inline V3 operator*(float f, V3 a) { return { a.x * f, a.y * f, a.z * f }; }
inline V3 operator/(V3 a, float f) { return { a.x / f, a.y / f, a.z / f }; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline float norm2(V3 a) { return dot(a, a); }
inline V3 normalized(V3 a) // Eigen normalized(): v / sqrt(squaredNorm) when > 0
{
	const float z = norm2(a);
	return z > 0 ? a / std::sqrt(z) : a;
}
inline bool isZero(V3 a, float prec) { return std::abs(a.x) <= prec && std::abs(a.y) <= prec && std::abs(a.z) <= prec; }
inline float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct Blob { // SpectralBlob, src/core/spectral/SpectralBlob.h:7-20
	float v[4];
	float& operator[](int i) { return v[i]; }
	float operator[](int i) const { return v[i]; }
};
inline Blob blob(float f) { return Blob{ { f, f, f, f } }; }
inline Blob operator*(Blob a, Blob b) { return { { a[0] * b[0], a[1] * b[1], a[2] * b[2], a[3] * b[3] } }; }
inline Blob operator*(Blob a, float f) { return { { a[0] * f, a[1] * f, a[2] * f, a[3] * f } }; }
inline Blob operator/(Blob a, Blob b) { return { { a[0] / b[0], a[1] / b[1], a[2] / b[2], a[3] / b[3] } }; }
inline Blob operator/(Blob a, float f) { return { { a[0] / f, a[1] / f, a[2] / f, a[3] / f } }; }
inline Blob operator+(Blob a, Blob b) { return { { a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3] } }; }
inline float bsum(Blob a) { return ((a[0] + a[1]) + a[2]) + a[3]; }
inline Blob heroOnly() { return Blob{ { 1, 0, 0, 0 } }; } // SpectralBlobUtils::HeroOnly
inline bool allLE(Blob a, float e) { return a[0] <= e && a[1] <= e && a[2] <= e && a[3] <= e; }
inline bool blobIsZero(Blob a, float e) { return std::abs(a[0]) <= e && std::abs(a[1]) <= e && std::abs(a[2]) <= e && std::abs(a[3]) <= e; }

// src/base/config/MathGlue.inl:8-24
inline float diffProd(float a, float b, float c, float d)
{
	const float cd	= c * d;
	const float err = std::fma(-c, d, cd);
	const float dop = std::fma(a, b, -cd);
	return dop + err;
}
inline float sumProd(float a, float b, float c, float d) { return std::fma(a, b, c * d); }

// ---------------------------------------------------------------- ShadingVector, src/base/math/ShadingVector.h
inline float cosTheta(V3 v) { return v.z; }
inline float cos2Theta(V3 v) { return v.z * v.z; }
inline float absCosTheta(V3 v) { return std::abs(v.z); }
inline float sin2Theta(V3 v) { return std::max(0.0f, 1 - cos2Theta(v)); }
inline float sinTheta(V3 v) { return std::sqrt(sin2Theta(v)); }
inline float tan2Theta(V3 v) { return absCosTheta(v) <= PR_EPSILON ? 0 : sin2Theta(v) / cos2Theta(v); }
inline float cos2Phi(V3 v)
{
	const float s = sin2Theta(v);
	return s <= PR_EPSILON ? 0 : std::min(1.0f, v.x * v.x / s);
}
inline float sin2Phi(V3 v)
{
	const float s = sin2Theta(v);
	return s <= PR_EPSILON ? 0 : std::min(1.0f, v.y * v.y / s);
}
inline bool sameHemisphere(V3 a, V3 b) { return std::signbit(a.z) == std::signbit(b.z); }
inline bool isPositiveHemisphere(V3 a) { return !std::signbit(a.z); }
inline V3 makeSameHemisphere(V3 self, V3 other) { return sameHemisphere(self, other) ? other : -other; }
inline V3 makePositiveHemisphere(V3 a) { return isPositiveHemisphere(a) ? a : -a; }

// ---------------------------------------------------------------- transcendental functions
// The reference calls std::sin/cos/tan/atan/atan2/acos on float = the libm it links; glibc's float versions are within
// 1 ulp but not correctly rounded (glibc 2.39: 1.3 % of sinf, 7.7 % of acosf, 15 % of atan2f results differ from the
// correctly rounded value), so the reference's last bit is libm-version dependent.  The oracle (and the device code,
// dev_math.cuh) use the correctly rounded fp32 value: evaluate in fp64, round once.
inline float cr_sin(float x) { return (float)std::sin((double)x); }
inline float cr_cos(float x) { return (float)std::cos((double)x); }
inline float cr_tan(float x) { return (float)std::tan((double)x); }
inline float cr_atanh(float x) { return (float)std::atanh((double)x); }
inline float cr_cosh(float x) { return (float)std::cosh((double)x); }
inline float cr_atan(float x) { return (float)std::atan((double)x); }
inline float cr_atan2(float y, float x) { return (float)std::atan2((double)y, (double)x); }
inline float cr_acos(float x) { return (float)std::acos((double)x); }

// ---------------------------------------------------------------- Sampling, src/base/math/Sampling.h:38-57
inline V3 cos_hemi(float u1, float u2)
{
	const float cosT   = std::sqrt(u1);
	const float sinT   = std::sqrt(1 - u1);
	const float phi	   = 2 * PR_PI * u2;
	const float sinPhi = cr_sin(phi);
	const float cosPhi = cr_cos(phi);
	return mk(sinT * cosPhi, sinT * sinPhi, cosT);
}
inline float cos_hemi_pdf(float NdotL) { return NdotL * PR_INV_PI; }

// ---------------------------------------------------------------- Scattering, src/base/math/Scattering.h:49-183
inline float refraction_angle(float cosI, float eta)
{
	if (std::signbit(cosI))
		return refraction_angle(-cosI, 1 / eta);
	const float k = 1 - (eta * eta) * (1 - cosI * cosI);
	return k < 0 ? -1.0f : std::sqrt(k);
}
inline V3 reflect(V3 V) { return mk(-V.x, -V.y, V.z); }
inline V3 reflect(V3 V, V3 N) { return (2 * dot(N, V)) * N - V; }
inline V3 refract(float eta, V3 wIn)
{
	if (std::signbit(wIn.z))
		return -refract(1 / eta, -wIn);
	const float cosT = refraction_angle(wIn.z, eta);
	if (cosT < 0.0f)
		return reflect(wIn);
	return normalized(mk(-wIn.x * eta, -wIn.y * eta, -cosT));
}
inline V3 refract(float eta, V3 wIn, V3 N, bool& total)
{
	const float cosI = dot(wIn, N);
	if (std::signbit(cosI))
		return -refract(1 / eta, -wIn, N, total);
	const float cosT = refraction_angle(cosI, eta);
	total			 = cosT < 0.0f;
	if (total)
		return reflect(wIn, N);
	return normalized((-wIn) * eta + (eta * cosI - cosT) * N);
}
inline V3 halfway_reflection(V3 wIn, V3 wOut) { return normalized(wIn + wOut); }
inline V3 halfway_refractive(float n_in, V3 wIn, float n_out, V3 wOut) { return -normalized(n_in * wIn + n_out * wOut); }
inline float reflective_jacobian(float cosO)
{
	const float denom = 4 * std::abs(cosO);
	return denom <= PR_EPSILON ? 0.0f : 1 / denom;
}
inline float refractive_jacobian(float eta, float cosI, float cosO)
{
	const float denom  = eta * cosI + cosO;
	const float denom2 = denom * denom;
	return denom2 <= PR_EPSILON ? 0.0f : std::abs(cosO) / denom2;
}

// ---------------------------------------------------------------- Fresnel, src/base/math/Fresnel.h:9-77
inline float fresnel_dielectric4(float cosI, float cosO, float n_in, float n_out)
{
	const float perp = diffProd(n_in, cosI, n_out, cosO) / sumProd(n_in, cosI, n_out, cosO);
	const float para = diffProd(n_out, cosI, n_in, cosO) / sumProd(n_out, cosI, n_in, cosO);
	return std::min(std::max(sumProd(para, para, perp, perp) / 2.0f, 0.0f), 1.0f);
}
inline float fresnel_dielectric(float cosI, float n_in, float n_out)
{
	if (std::signbit(cosI))
		return fresnel_dielectric(-cosI, n_out, n_in);
	const float cosT = refraction_angle(cosI, n_in / n_out);
	if (cosT < 0)
		return 1;
	return fresnel_dielectric4(cosI, cosT, n_in, n_out);
}
inline float fresnel_conductor(float cosI, float n_in, float n_out, float k)
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
	const float ap	   = std::sqrt(sumProd(t0, t0, 4 * eta2, kappa2));
	const float t1	   = ap + cosI2;
	const float a	   = std::sqrt((ap + t0) / 2);
	const float t2	   = 2 * cosI * a;
	const float perp2  = (t1 - t2) / (t1 + t2);
	const float t3	   = sumProd(cosI2, ap, sinI2, sinI2);
	const float t4	   = t2 * sinI2;
	const float para2  = perp2 * (t3 - t4) / (t3 + t4);
	const float R	   = (para2 + perp2) / 2;
	return std::min(std::max(R, 0.0f), 1.0f);
}
inline float schlick_term(float d)
{
	const float t = 1 - d;
	return (t * t) * (t * t) * t;
}
inline float schlick(float d, float f0) { return f0 + (1 - f0) * schlick_term(d); }
inline float schlick3(float d, float n1, float n2)
{
	const float c = (n1 - n2) / (n1 + n2);
	return schlick(d, c * c);
}

// ---------------------------------------------------------------- Microfacet, src/base/math/Microfacet.h
inline float g_1_smith_opt(float NdotK, float roughness) // :56-62
{
	const float a	  = roughness * roughness;
	const float b	  = NdotK * NdotK;
	const float denom = NdotK + std::sqrt(a + b - a * b);
	return (denom <= PR_EPSILON) ? 0.0f : 1.0f / denom;
}
inline float g_1_smith(V3 K, float roughness) // :76-82
{
	const float a	  = roughness * roughness;
	const float b	  = tan2Theta(K);
	const float denom = 1 + std::sqrt(1 + a * b);
	return (denom <= PR_EPSILON) ? 0.0f : 2.0f / denom;
}
inline float g_1_smith(V3 K, float rx, float ry) // :83-90
{
	const float ax2	  = cos2Phi(K) * rx * rx;
	const float ay2	  = sin2Phi(K) * ry * ry;
	const float b	  = tan2Theta(K);
	const float denom = 1 + std::sqrt(1 + (ax2 + ay2) * b);
	return (denom <= PR_EPSILON) ? 0.0f : 2.0f / denom;
}
inline float g_1_smith_lambda(V3 K, float roughness) // :91-96
{
	const float a = roughness * roughness;
	const float b = tan2Theta(K);
	return (std::sqrt(1 + a * b) - 1) / 2;
}
inline float g_1_smith_lambda(V3 K, float rx, float ry) // :97-103
{
	const float ax2 = cos2Phi(K) * rx * rx;
	const float ay2 = sin2Phi(K) * ry * ry;
	const float b	= tan2Theta(K);
	return (std::sqrt(1 + (ax2 + ay2) * b) - 1) / 2;
}
inline float ndf_ggx(V3 H, float roughness) // :122-137
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
inline float ndf_ggx(V3 H, float rx, float ry) // :138-159
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
inline float pdf_ggx(V3 H, float r) { return ndf_ggx(H, r) * absCosTheta(H); }						// :213-216
inline float pdf_ggx(V3 H, float rx, float ry) { return ndf_ggx(H, rx, ry) * absCosTheta(H); } // :217-220
inline V3 spherical_cartesian(float thSin, float thCos, float phSin, float phCos) { return mk(thSin * phCos, thSin * phSin, thCos); }
inline V3 sample_ndf_ggx(float u0, float u1, float roughness) // :221-233
{
	const float alpha2 = roughness * roughness;
	const float t2	   = alpha2 * u1 / (1 - u1);
	const float cosT   = alpha2 <= PR_EPSILON ? 1.0f : std::max(0.001f, 1.0f / std::sqrt(1 + t2));
	const float sinT   = std::sqrt(1 - cosT * cosT);
	const float sinPhi = cr_sin(2 * PR_PI * u0);
	const float cosPhi = cr_cos(2 * PR_PI * u0);
	return spherical_cartesian(sinT, cosT, sinPhi, cosPhi);
}
inline V3 sample_ndf_ggx(float u0, float u1, float rx, float ry) // :234-249
{
	const float phi	   = cr_atan(ry / rx * cr_tan(PR_PI + 2 * PR_PI * u0)) + PR_PI * std::floor(2 * u0 + 0.5f);
	const float sinPhi = cr_sin(phi);
	const float cosPhi = cr_cos(phi);
	const float f1	   = cosPhi / rx;
	const float f2	   = sinPhi / ry;
	const float alpha2 = 1 / (f1 * f1 + f2 * f2);
	const float t2	   = alpha2 * u1 / (1 - u1);
	const float cosT   = std::max(0.001f, 1.0f / std::sqrt(1 + t2));
	const float sinT   = std::sqrt(1 - cosT * cosT);
	return spherical_cartesian(sinT, cosT, sinPhi, cosPhi);
}
inline float pdf_ggx_vndf(V3 V, V3 H, float rx, float ry) // :253-259
{
	return absCosTheta(V) <= PR_EPSILON ? 0.0f : g_1_smith(V, rx, ry) * std::abs(dot(V, H)) * ndf_ggx(H, rx, ry) / absCosTheta(V);
}
inline V3 sample_vndf_ggx(float u0, float u1, V3 nV, float rx, float ry) // :261-330 (the #if 1 branch)
{
	const V3 Vh		  = normalized(mk(rx * nV.x, ry * nV.y, nV.z));
	const float lensq = sumProd(Vh.x, Vh.x, Vh.y, Vh.y);
	const V3 T1		  = lensq > PR_EPSILON ? mk(-Vh.y, Vh.x, 0) / std::sqrt(lensq) : mk(1, 0, 0);
	const V3 T2		  = cross(Vh, T1);
	const float r	  = std::sqrt(u0);
	const float phi	  = 2.0f * PR_PI * u1;
	const float t1	  = r * cr_cos(phi);
	float t2		  = r * cr_sin(phi);
	const float s	  = 0.5f * (1.0f + Vh.z);
	t2				  = (1.0f - s) * std::sqrt(1.0f - t1 * t1) + s * t2;
	const V3 Nh		  = t1 * T1 + t2 * T2 + std::sqrt(std::max(0.0f, 1.0f + diffProd(-t1, t1, t2, t2))) * Vh;
	return normalized(mk(rx * Nh.x, ry * Nh.y, std::max(0.0f, Nh.z)));
}

// ---------------------------------------------------------------- RoughDistribution, src/base/math/RoughDistribution.h
struct RoughDistribution {
	float M1, M2;
	bool aniso, vndf;
	bool isDelta() const { return M1 <= 1e-3f || M2 <= 1e-3f; }
	float G(V3 H, V3 V, V3 L) const
	{
		const bool chi_v = cosTheta(V) * dot(H, V) > PR_EPSILON;
		const bool chi_l = cosTheta(L) * dot(H, L) > PR_EPSILON;
		if (!chi_v || !chi_l)
			return 0.0f;
		if (!vndf)
			return aniso ? g_1_smith(V, M1, M2) * g_1_smith(L, M1, M2) : g_1_smith(V, M1) * g_1_smith(L, M1);
		const float denom = aniso ? 1 + g_1_smith_lambda(V, M1, M2) + g_1_smith_lambda(L, M1, M2) : 1 + g_1_smith_lambda(V, M1) + g_1_smith_lambda(L, M1);
		return denom <= PR_EPSILON ? 0.0f : 1 / denom;
	}
	float D(V3 H) const { return aniso ? ndf_ggx(H, M1, M2) : ndf_ggx(H, M1); }
	float Norm(V3 H, V3 V, V3 L) const
	{
		const float denom = absCosTheta(V);
		if (denom <= PR_EPSILON)
			return 0;
		return std::abs(dot(H, L)) / denom;
	}
	float DGNorm(V3 H, V3 V, V3 L) const { return D(H) * G(H, V, L) * Norm(H, V, L); }
	float pdf(V3 H, V3 V) const
	{
		if (isDelta())
			return 1.0f;
		if (vndf)
			return pdf_ggx_vndf(makePositiveHemisphere(V), makePositiveHemisphere(H), M1, M2);
		return aniso ? pdf_ggx(H, M1, M2) : pdf_ggx(H, M1);
	}
	V3 sample(float r0, float r1, V3 V) const
	{
		if (isDelta())
			return mk(0, 0, 1);
		if (vndf)
			return sample_vndf_ggx(r0, r1, makePositiveHemisphere(V), M1, M2);
		return aniso ? sample_ndf_ggx(r0, r1, M1, M2) : sample_ndf_ggx(r0, r1, M1);
	}
};

// ---------------------------------------------------------------- MicrofacetReflection.h
struct MicrofacetReflection {
	RoughDistribution D;
	bool isDelta() const { return D.isDelta(); }
	float evalDielectric(V3 wIn, V3 wOut, float n_in, float n_out) const
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
	float evalConductor(V3 wIn, V3 wOut, float ior, float kappa) const
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
	float eval(V3 wIn, V3 wOut) const
	{
		if (!sameHemisphere(wIn, wOut))
			return 0.0f;
		const V3 H = halfway_reflection(wIn, wOut);
		if (isDelta())
			return 1.0f;
		const float cosI = dot(H, wIn);
		return D.DGNorm(H, wIn, wOut) * reflective_jacobian(cosI);
	}
	float pdf(V3 wIn, V3 wOut) const
	{
		if (!sameHemisphere(wIn, wOut))
			return 0.0f;
		const V3 H = halfway_reflection(wIn, wOut);
		if (isDelta())
			return 1.0f;
		const float cosI = dot(H, wIn);
		return reflective_jacobian(cosI) * D.pdf(H, wIn);
	}
	V3 sample(float r0, float r1, V3 wIn) const
	{
		const V3 H = D.sample(r0, r1, wIn);
		if (isZero(H, PR_EPSILON))
			return mk(0, 0, 0);
		const V3 wOut = reflect(wIn, H);
		return sameHemisphere(wIn, wOut) ? wOut : mk(0, 0, 0);
	}
};

// ---------------------------------------------------------------- MicrofacetTransmission.h
struct MicrofacetTransmission {
	RoughDistribution D;
	float InnerIOR, OuterIOR;
	bool isDelta() const { return D.isDelta(); }
	bool setup(V3 wIn, V3 wOut, V3& H, float& cosI, float& cosO, float& eta) const
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
	float evalDielectric(V3 wIn, V3 wOut, bool isLightPath) const
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
	float pdf(V3 wIn, V3 wOut) const
	{
		V3 H;
		float cosI, cosO, eta;
		if (!setup(wIn, wOut, H, cosI, cosO, eta))
			return 0.0f;
		if (isDelta())
			return 1.0f;
		return D.pdf(H, wIn) * refractive_jacobian(eta, cosI, cosO);
	}
	V3 sample(float r0, float r1, V3 wIn) const
	{
		const V3 H = D.sample(r0, r1, wIn);
		if (isZero(H, PR_EPSILON))
			return mk(0, 0, 0);
		const float eta = InnerIOR / OuterIOR;
		bool total;
		const V3 L = refract(eta, wIn, H, total);
		return (total == sameHemisphere(wIn, L)) ? L : mk(0, 0, 0);
	}
};

// ---------------------------------------------------------------- Tangent.h:9-56, Transform.h:9-32
inline V3 fromTangentSpace(V3 N, V3 Nx, V3 Ny, V3 V) { return normalized((N * V.z + Ny * V.y) + Nx * V.x); }
inline V3 toTangentSpace(V3 N, V3 Nx, V3 Ny, V3 V) { return normalized(mk(dot(Nx, V), dot(Ny, V), dot(N, V))); }
inline void frame_duff(V3 N, V3& Nx, V3& Ny)
{
	const float sign = copysignf(1.0f, N.z);
	const float a	 = -1.0f / (sign + N.z);
	const float b	 = N.x * N.y * a;
	Nx				 = mk(1.0f + sign * N.x * N.x * a, sign * b, -sign * N.x);
	Ny				 = mk(b, sign + N.y * N.y * a, -N.y);
}
inline void tangent_frame(V3 N, V3& Nx, V3& Ny)
{
	frame_duff(N, Nx, Ny);
	Nx = normalized(Nx);
	Ny = normalized(Ny);
}
inline float nextFloatUp(float v)
{ // src/base/math/Bits.h nextFloatUp
	if (std::isinf(v) && v > 0.0f)
		return v;
	if (v == -0.0f)
		v = 0.0f;
	uint32_t ui;
	std::memcpy(&ui, &v, 4);
	if (v >= 0)
		++ui;
	else
		--ui;
	std::memcpy(&v, &ui, 4);
	return v;
}
inline float nextFloatDown(float v)
{
	if (std::isinf(v) && v < 0.0f)
		return v;
	if (v == 0.0f)
		v = -0.0f;
	uint32_t ui;
	std::memcpy(&ui, &v, 4);
	if (v > 0)
		--ui;
	else
		++ui;
	std::memcpy(&v, &ui, 4);
	return v;
}
inline V3 safePosition(V3 pos, V3 dir, V3 N)
{
	const float d = ((std::abs(N.x) * 0.0001f + std::abs(N.y) * 0.0001f) + std::abs(N.z) * 0.0001f);
	V3 offset	  = d * N;
	if (dot(dir, N) < 0)
		offset = -offset;
	V3 p	  = pos + offset;
	float* pp = &p.x;
	const float* oo = &offset.x;
	for (int i = 0; i < 3; ++i) {
		if (oo[i] > 0)
			pp[i] = nextFloatUp(pp[i]);
		else if (oo[i] < 0)
			pp[i] = nextFloatDown(pp[i]);
	}
	return p;
}

// ---------------------------------------------------------------- Spherical.h:9-53
inline void spherical_from_direction(V3 D, float& theta, float& phi)
{
	const float x = (D.x == 0 && D.y == 0) ? 1e-5f : D.x;
	phi			  = cr_atan2(D.y, x);
	phi			  = phi < 0 ? phi + 2 * PR_PI : phi;
	theta		  = cr_acos(D.z);
}
inline void uv_from_normal(V3 N, float& u, float& v)
{
	float theta, phi;
	spherical_from_direction(N, theta, phi);
	const float tx = theta * PR_INV_PI, ty = phi * PR_INV_PI;
	u = ty / 2;
	v = tx;
}
inline V3 cartesian_from_uv(float u, float v)
{
	const float theta = v * PR_PI, phi = u * 2 * PR_PI;
	return spherical_cartesian(cr_sin(theta), cr_cos(theta), cr_sin(phi), cr_cos(phi));
}
} // namespace orc
