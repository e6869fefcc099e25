// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_math.h).  PARITY PINNING: the math layer is pinned against
// the reference's own known-answer tests (tests/test_oracle_*.py port src/tests/{fresnel,microfacets,scattering,
// sampling,tangent,materials,random}.cpp); the intersection stage restates Embree 3's documented robust kernels
// (Embree is an un-vendored dependency of the reference and absent here) -> "parity unpinned" for hit ids
// against Embree itself; see DESIGN.md.
//
// CPU restatement of the PearRay spectral path tracer ('direct' integrator) consuming the same POD scene
// descriptor (include/prb200_abi.h) as the CUDA library.  Depth-first, one path at a time, exactly in the order
// of the reference (src/plugins/main/integrators/direct.cpp + src/vcm/vcm/Walker.h), scalar fp32, FTZ/DAZ.
#include "../include/prb200_abi.h"
#include "oracle_math.h"

#include <atomic>
#include <cstring>
#include <thread>
#include <vector>
#include <xmmintrin.h>
#include <pmmintrin.h>

using namespace orc;

namespace {
// ------------------------------------------------------------------ RNG: src/core/Random.h:26-179
struct Rng {
	uint64_t s;
	uint32_t get32()
	{ // pcg32_fast = mcg_xsh_rs_64_32, src/core/random/pcg_random.hpp:812-836,1865
		const uint64_t old = s;
		s				   = old * 6364136223846793005ULL;
		const uint32_t rs  = (uint32_t)(old >> 61);
		const uint64_t x   = old ^ (old >> 22);
		return (uint32_t)(x >> (22 + rs));
	}
	float getFloat()
	{ // Random.h:133-158
		const uint32_t u = (get32() >> 9) | 0x3F800000u;
		float f;
		std::memcpy(&f, &u, 4);
		return f - 1.0f;
	}
	// Vector2f(getFloat(), getFloat()): unsequenced in the reference; GCC evaluates right-to-left, so the FIRST
	// draw becomes y and the second x (same convention in host and device code)
	void get2D(float& x, float& y)
	{
		y = getFloat();
		x = getFloat();
	}
};

struct Scene {
	const prb_scene_desc* d;
	std::vector<float> rrProb; // RussianRoulette::probability table
};

inline V3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }
inline V3 xfPoint(const float* m, V3 p) { return mk(((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3], ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7], ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]); }
inline V3 xfVec(const float* m, V3 p) { return mk((m[0] * p.x + m[1] * p.y) + m[2] * p.z, (m[4] * p.x + m[5] * p.y) + m[6] * p.z, (m[8] * p.x + m[9] * p.y) + m[10] * p.z); }
inline V3 m3mul(const float* m, V3 p) { return mk((m[0] * p.x + m[1] * p.y) + m[2] * p.z, (m[3] * p.x + m[4] * p.y) + m[5] * p.z, (m[6] * p.x + m[7] * p.y) + m[8] * p.z); }

// ------------------------------------------------------------------ spectra / nodes
// EquidistantSpectrumView::lookup, src/core/spectral/EquidistantSpectrum.inl:34-41
inline float tableLookup(const float* data, uint32_t count, float start, float end, float w)
{
	const float delta = (end - start) / (count - 1);
	const float af	  = std::max(0.0f, (w - start) / delta);
	const int index	  = (int)std::min<float>((float)(count - 2), af);
	const float t	  = std::min<float>((float)(count - 1), af) - index;
	return data[index] * (1 - t) + data[index + 1] * t;
}
constexpr float CIE_START = 390, CIE_END = 830, CIE_RANGE = CIE_END - CIE_START; // src/core/spectral/CIE.h:18-29
constexpr int CIE_N			 = 441;
constexpr float CIE_Y_NORM = 113.042314572337f * (CIE_RANGE / (CIE_N - 1));
inline float cieEval(const Scene& sc, int c, float w)
{ // CIE::eval_x/y/z, CIE.h:41-58
	return tableLookup(sc.d->pool + sc.d->cie_offset + c * CIE_N, CIE_N, CIE_START, CIE_END, w) / CIE_Y_NORM * CIE_RANGE;
}

// ---- image textures: NonParametricImageNode::eval, loader/shader/ImageNode.cpp:131-162.  OpenImageIO's texture system is an
// un-vendored dependency of the reference; the lookup restates its documented behaviour without MIP levels and derivatives
// (texel centres at (i + 0.5) / size, wrap modes black / clamp / periodic / mirror, closest / bilinear / B-spline bicubic).
bool wrapTexel(int& i, int size, int mode)
{
	if (i >= 0 && i < size)
		return true;
	switch (mode) {
	default: return false; // black
	case PRB_WRAP_CLAMP: i = i < 0 ? 0 : size - 1; return true;
	case PRB_WRAP_PERIODIC:
		i %= size;
		if (i < 0)
			i += size;
		return true;
	case PRB_WRAP_MIRROR: {
		const int period = 2 * size;
		i %= period;
		if (i < 0)
			i += period;
		if (i >= size)
			i = period - 1 - i;
		return true;
	}
	}
}
void fetchTexel(const float* img, int w, int h, int x, int y, int wrapS, int wrapT, float rgb[3])
{
	rgb[0] = rgb[1] = rgb[2] = 0;
	if (!wrapTexel(x, w, wrapS) || !wrapTexel(y, h, wrapT))
		return;
	const float* t = img + 3 * ((size_t)y * w + x);
	rgb[0] = t[0], rgb[1] = t[1], rgb[2] = t[2];
}
void bsplineWeights(float f, float w[4])
{
	const float one_f = 1.0f - f;
	w[0]			  = (one_f * one_f * one_f) / 6.0f;
	w[1]			  = 2.0f / 3.0f - 0.5f * f * f * (2.0f - f);
	w[2]			  = 2.0f / 3.0f - 0.5f * one_f * one_f * (2.0f - one_f);
	w[3]			  = (f * f * f) / 6.0f;
}
void upsamplerPrepare(const Scene& sc, const float rgb[3], float coeffs[3])
{ // SpectralUpsampler::prepare for one triple, src/core/spectral/SpectralUpsampler.cpp
	constexpr float EPS = 0.0001f;
	if (rgb[0] <= EPS && rgb[1] <= EPS && rgb[2] <= EPS) {
		coeffs[0] = 0, coeffs[1] = 0, coeffs[2] = -500.0f;
		return;
	}
	if (1 - rgb[0] <= EPS && 1 - rgb[1] <= EPS && 1 - rgb[2] <= EPS) {
		coeffs[0] = 0, coeffs[1] = 0, coeffs[2] = 5000000.0f;
		return;
	}
	const uint32_t res = sc.d->upsampler_res;
	const float* scale = sc.d->pool + sc.d->upsampler_offset;
	const float* d	   = scale + res;
	const uint32_t dx = 3, dy = 3 * res, dz = 3 * res * res;
	int largest = 0;
	for (int j = 1; j < 3; ++j)
		if (rgb[largest] <= rgb[j])
			largest = j;
	const float z	  = rgb[largest];
	const float scl	  = (float)(res - 1) / z;
	const float x	  = rgb[(largest + 1) % 3] * scl;
	const float y	  = rgb[(largest + 2) % 3] * scl;
	const uint32_t xi = std::min((uint32_t)x, res - 2);
	const uint32_t yi = std::min((uint32_t)y, res - 2);
	int left = 0, size = (int)res - 2;
	const int lastInterval = (int)res - 2;
	while (size > 0) {
		const int half = size >> 1, middle = left + half + 1;
		if (scale[middle] < z) {
			left = middle;
			size -= half + 1;
		} else {
			size = half;
		}
	}
	const uint32_t zi = (uint32_t)std::min(left, lastInterval);
	uint32_t off	  = (((largest * res + zi) * res + yi) * res + xi) * 3;
	const float x1 = x - (float)xi, x0 = 1.0f - x1, y1 = y - (float)yi, y0 = 1.0f - y1;
	const float z1 = (z - scale[zi]) / (scale[zi + 1] - scale[zi]), z0 = 1.0f - z1;
	for (int j = 0; j < 3; ++j) {
		coeffs[j] = ((d[off] * x0 + d[off + dx] * x1) * y0 + (d[off + dy] * x0 + d[off + dy + dx] * x1) * y1) * z0
					+ ((d[off + dz] * x0 + d[off + dz + dx] * x1) * y0 + (d[off + dz + dy] * x0 + d[off + dz + dy + dx] * x1) * y1) * z1;
		++off;
	}
}
float srgbLinearize(float x)
{ // RGBConverter::linearize, src/core/spectral/RGBConverter.cpp:53-58 -- as written there, `x / 12.92 * x` on the linear segment
	if (x <= 0.04045f)
		return x / 12.92f * x;
	return (float)std::pow((double)((x + 0.055f) / 1.055f), (double)2.4f);
}
Blob evalImageNode(const Scene& sc, const prb_node& n, const Blob& wvl, float u, float v)
{
	const int w = (int)(n.b & 0xFFFFu), h = (int)(n.b >> 16);
	const float* img = sc.d->pool + n.a;
	const int interp = (int)n.p[0], wrapS = (int)n.p[1], wrapT = (int)n.p[2];
	const float x = u * (float)w - 0.5f, y = (1 - v) * (float)h - 0.5f; // texture(s = u, t = 1 - v), ImageNode.cpp:141-145
	const float flx = std::floor(x), fly = std::floor(y);
	int ix = (int)flx, iy = (int)fly;
	const float fx = x - flx, fy = y - fly;
	float rgb[3];
	if (interp == PRB_TEX_CLOSEST) {
		if (fx > 0.5f)
			++ix;
		if (fy > 0.5f)
			++iy;
		fetchTexel(img, w, h, ix, iy, wrapS, wrapT, rgb);
	} else if (interp == PRB_TEX_BILINEAR) {
		float c00[3], c10[3], c01[3], c11[3];
		fetchTexel(img, w, h, ix, iy, wrapS, wrapT, c00);
		fetchTexel(img, w, h, ix + 1, iy, wrapS, wrapT, c10);
		fetchTexel(img, w, h, ix, iy + 1, wrapS, wrapT, c01);
		fetchTexel(img, w, h, ix + 1, iy + 1, wrapS, wrapT, c11);
		for (int c = 0; c < 3; ++c)
			rgb[c] = (c00[c] * (1 - fx) + c10[c] * fx) * (1 - fy) + (c01[c] * (1 - fx) + c11[c] * fx) * fy;
	} else {
		float wx[4], wy[4];
		bsplineWeights(fx, wx);
		bsplineWeights(fy, wy);
		rgb[0] = rgb[1] = rgb[2] = 0;
		for (int j = 0; j < 4; ++j) {
			float row[3] = { 0, 0, 0 };
			for (int i = 0; i < 4; ++i) {
				float t[3];
				fetchTexel(img, w, h, ix - 1 + i, iy - 1 + j, wrapS, wrapT, t);
				for (int c = 0; c < 3; ++c)
					row[c] += wx[i] * t[c];
			}
			for (int c = 0; c < 3; ++c)
				rgb[c] += wy[j] * row[c];
		}
	}
	if (n.p[3] != 0.0f)
		for (int c = 0; c < 3; ++c)
			rgb[c] = srgbLinearize(rgb[c]);
	float k[3];
	upsamplerPrepare(sc, rgb, k);
	Blob r;
	for (int i = 0; i < 4; ++i) { // SpectralUpsampler::compute, SpectralUpsampler.h:45-49
		const float q = (k[0] * wvl[i] + k[1]) * wvl[i] + k[2];
		r[i]		  = 0.5f * q * (1.0f / std::sqrt(q * q + 1.0f)) + 0.5f;
	}
	return r;
}

Blob evalNode(const Scene& sc, uint32_t id, const Blob& w, float u, float v)
{
	const prb_node& n = sc.d->nodes[id];
	Blob r;
	switch (n.type) {
	case PRB_NODE_IMAGE: return evalImageNode(sc, n, w, u, v);
	default:
	case PRB_NODE_CONST: return blob(n.p[0]);
	case PRB_NODE_PARAM:
	case PRB_NODE_PARAM_SCALED: // SpectralUpsampler::compute, src/core/spectral/SpectralUpsampler.h:45-49
		for (int i = 0; i < 4; ++i) {
			const float x = (n.p[0] * w[i] + n.p[1]) * w[i] + n.p[2];
			r[i]		  = 0.5f * x * (1.0f / std::sqrt(x * x + 1.0f)) + 0.5f;
			if (n.type == PRB_NODE_PARAM_SCALED)
				r[i] = r[i] * n.p[3];
		}
		return r;
	case PRB_NODE_TABLE:
		for (int i = 0; i < 4; ++i)
			r[i] = tableLookup(sc.d->pool + n.a, n.b, n.p[0], n.p[1], w[i]);
		return r;
	case PRB_NODE_SELLMEIER: { // Scattering::sellmeier2 + sqrt, src/base/math/Scattering.h:219-242
		const float* B = sc.d->pool + n.a;
		const float* C = B + n.b;
		for (int i = 0; i < 4; ++i) {
			const float qm	= w[i] / 1000;
			const float qm2 = qm * qm;
			float value		= 1;
			for (uint32_t k = 0; k < n.b; ++k)
				value += B[k] * qm2 / (qm2 - C[k]);
			r[i] = std::sqrt(value);
		}
		return r;
	}
	case PRB_NODE_MUL: return evalNode(sc, n.a, w, u, v) * evalNode(sc, n.b, w, u, v);
	case PRB_NODE_CHECKER: { // CheckerboardNode.cpp:26-48
		float cu = u, cv = v;
		if (n.p[2] == 1.0f) {
			cu = u * n.p[0];
			cv = v * n.p[0];
		} else if (n.p[2] == 2.0f) {
			cu = u * n.p[0];
			cv = v * n.p[1];
		}
		const bool check = ((int)std::floor(cu) + (int)std::floor(cv)) % 2 == 0;
		return check ? evalNode(sc, n.b, w, u, v) : evalNode(sc, n.a, w, u, v);
	}
	}
}

// ------------------------------------------------------------------ intersection
struct Hit {
	uint32_t entity = PRB_INVALID_ID, prim = 0;
	float u = 0, v = 0, t = PR_INF;
};
inline bool better(float t, uint32_t e, uint32_t p, const Hit& h)
{ // order independent closest-hit rule: lexicographic (t, entity, prim)
	if (h.entity == PRB_INVALID_ID)
		return true;
	if (t != h.t)
		return t < h.t;
	if (e != h.entity)
		return e < h.entity;
	return p < h.prim;
}

// Embree 3 PlueckerIntersector (kernels/geometry/triangle_intersector_pluecker.h), scalar restatement:
// edge functions of the ray against the triangle edges relative to the ray origin, accepted when all have
// the same sign within ulp*|U+V+W|; t = (v0.Ng)/(d.Ng) with the "stable" normal; two-sided.
// Embree writes its vector algebra with msub / madd (common/math/vec3.h: cross = msub(a.y, b.z, a.z * b.y) ..., dot =
// madd(a.x, b.x, madd(a.y, b.y, a.z * b.z))), which ARE fused multiply-adds in the AVX2 / AVX-512 kernels its ISA dispatch
// selects on any current host, so the restatement uses explicit std::fma in exactly those places (the build adds -mfma).
inline float msub(float a, float b, float c) { return std::fma(a, b, -c); }
inline V3 crossE(V3 a, V3 b) { return mk(msub(a.y, b.z, a.z * b.y), msub(a.z, b.x, a.x * b.z), msub(a.x, b.y, a.y * b.x)); }
inline float dotE(V3 a, V3 b) { return std::fma(a.x, b.x, std::fma(a.y, b.y, a.z * b.z)); }
inline V3 stableTriangleNormal(V3 a, V3 b, V3 c)
{
	const float ab_x = a.z * b.y, ab_y = a.x * b.z, ab_z = a.y * b.x;
	const float bc_x = b.z * c.y, bc_y = b.x * c.z, bc_z = b.y * c.x;
	const V3 cross_ab = mk(msub(a.y, b.z, ab_x), msub(a.z, b.x, ab_y), msub(a.x, b.y, ab_z));
	const V3 cross_bc = mk(msub(b.y, c.z, bc_x), msub(b.z, c.x, bc_y), msub(b.x, c.y, bc_z));
	const bool sx = std::abs(ab_x) < std::abs(bc_x), sy = std::abs(ab_y) < std::abs(bc_y), sz = std::abs(ab_z) < std::abs(bc_z);
	return mk(sx ? cross_ab.x : cross_bc.x, sy ? cross_ab.y : cross_bc.y, sz ? cross_ab.z : cross_bc.z);
}
inline bool triTest(V3 O, V3 D, float tmin, float tmax, V3 p0, V3 p1, V3 p2, float& t, float& u, float& v)
{
	const V3 v0 = p0 - O, v1 = p1 - O, v2 = p2 - O;
	const V3 e0 = v2 - v0, e1 = v0 - v1, e2 = v1 - v2;
	const float U	= dotE(crossE(e0, v2 + v0), D);
	const float V	= dotE(crossE(e1, v0 + v1), D);
	const float W	= dotE(crossE(e2, v1 + v2), D);
	const float UVW = (U + V) + W;
	const float eps = PR_EPSILON * std::abs(UVW);
	const float mn = std::min(U, std::min(V, W)), mx = std::max(U, std::max(V, W));
	if (!(mn >= -eps || mx <= eps))
		return false;
	const V3 Ng		= stableTriangleNormal(e0, e1, e2);
	const float den = 2 * dotE(Ng, D);
	if (den == 0)
		return false;
	const float T = 2 * dotE(v0, Ng);
	t			  = T / den; // (Embree multiplies by a Newton-refined reciprocal here: not reproducible bit for bit)
	if (!(tmin <= t && t <= tmax))
		return false;
	if (UVW == 0) { // degenerate (edge-on / zero-area) configuration: Embree masks rcp(0) to 0
		u = 0;
		v = 0;
	} else {
		u = std::min(U / UVW, 1.0f);
		v = std::min(V / UVW, 1.0f);
	}
	return true;
}
// instance transform of the ray (Embree xfmPoint / xfmVector, common/math/affinespace.h: madd chains)
inline V3 xfPointE(const float* m, V3 p)
{
	return mk(std::fma(p.x, m[0], std::fma(p.y, m[1], std::fma(p.z, m[2], m[3]))), std::fma(p.x, m[4], std::fma(p.y, m[5], std::fma(p.z, m[6], m[7]))),
			  std::fma(p.x, m[8], std::fma(p.y, m[9], std::fma(p.z, m[10], m[11]))));
}
inline V3 xfVecE(const float* m, V3 p)
{
	return mk(std::fma(p.x, m[0], std::fma(p.y, m[1], p.z * m[2])), std::fma(p.x, m[4], std::fma(p.y, m[5], p.z * m[6])), std::fma(p.x, m[8], std::fma(p.y, m[9], p.z * m[10])));
}
// Embree 3 SphereIntersector1 (kernels/geometry/sphere_intersector.h), front hit first then back hit
inline bool sphereTest(V3 O, V3 D, float tmin, float tmax, V3 center, float radius, float& t)
{
	const float rd2	   = 1.0f / dot(D, D);
	const V3 c0		   = center - O;
	const float projC0 = dot(c0, D) * rd2;
	const V3 perp	   = c0 - projC0 * D;
	const float l2	   = dot(perp, perp);
	const float r2	   = radius * radius;
	if (!(l2 <= r2))
		return false;
	const float td		= std::sqrt((r2 - l2) * rd2);
	const float t_front = projC0 - td, t_back = projC0 + td;
	if (tmin <= t_front && t_front <= tmax) {
		t = t_front;
		return true;
	}
	if (tmin <= t_back && t_back <= tmax) {
		t = t_back;
		return true;
	}
	return false;
}

// per-mesh acceleration for the oracle: a plain median-split binary BVH (independent of the product's BVH8)
struct OBox {
	float lo[3], hi[3];
};
struct ONode {
	OBox box;
	uint32_t left, right, first, count; // count>0 => leaf
};
struct OMeshAccel {
	std::vector<ONode> nodes;
	std::vector<uint32_t> tris; // indices into the triangle list
};
struct OTri {
	V3 a, b, c;
	uint32_t prim, flags;
};
struct Accel {
	std::vector<std::vector<OTri>> meshTris;
	std::vector<OMeshAccel> meshAccel;
	std::vector<OBox> entityBox;
};

void faceTris(std::vector<OTri>& out, V3 v0, V3 v1, V3 v2, const V3* v3, uint32_t prim)
{ // Embree quads: (v0,v1,v3) and (v2,v3,v1), second flagged
	if (!v3) {
		out.push_back({ v0, v1, v2, prim, 0 });
	} else {
		out.push_back({ v0, v1, *v3, prim, 0 });
		out.push_back({ v2, *v3, v1, prim, 1 });
	}
}
OBox triBox(const OTri& t)
{
	OBox b;
	const float* p[3] = { &t.a.x, &t.b.x, &t.c.x };
	for (int k = 0; k < 3; ++k) {
		b.lo[k] = std::min(p[0][k], std::min(p[1][k], p[2][k]));
		b.hi[k] = std::max(p[0][k], std::max(p[1][k], p[2][k]));
		float m = std::max(std::abs(b.lo[k]), std::abs(b.hi[k]));
		m		= std::max(m, b.hi[k] - b.lo[k]);
		const float pad = std::max(4e-6f * m, 1e-30f);
		b.lo[k] -= pad;
		b.hi[k] += pad;
	}
	return b;
}
void buildNode(OMeshAccel& acc, const std::vector<OTri>& tris, const std::vector<OBox>& boxes, uint32_t node, uint32_t first, uint32_t count)
{
	OBox b;
	for (int k = 0; k < 3; ++k) {
		b.lo[k] = PR_INF;
		b.hi[k] = -PR_INF;
	}
	for (uint32_t i = first; i < first + count; ++i)
		for (int k = 0; k < 3; ++k) {
			b.lo[k] = std::min(b.lo[k], boxes[acc.tris[i]].lo[k]);
			b.hi[k] = std::max(b.hi[k], boxes[acc.tris[i]].hi[k]);
		}
	acc.nodes[node].box = b;
	if (count <= 4) {
		acc.nodes[node].first = first;
		acc.nodes[node].count = count;
		return;
	}
	int axis = 0;
	float ext = -1;
	for (int k = 0; k < 3; ++k)
		if (b.hi[k] - b.lo[k] > ext) {
			ext	 = b.hi[k] - b.lo[k];
			axis = k;
		}
	const uint32_t mid = first + count / 2;
	std::nth_element(acc.tris.begin() + first, acc.tris.begin() + mid, acc.tris.begin() + first + count, [&](uint32_t x, uint32_t y) {
		return boxes[x].lo[axis] + boxes[x].hi[axis] < boxes[y].lo[axis] + boxes[y].hi[axis];
	});
	const uint32_t l = (uint32_t)acc.nodes.size();
	acc.nodes.push_back(ONode{});
	acc.nodes.push_back(ONode{});
	acc.nodes[node].left  = l;
	acc.nodes[node].right = l + 1;
	acc.nodes[node].count = 0;
	buildNode(acc, tris, boxes, l, first, mid - first);
	buildNode(acc, tris, boxes, l + 1, mid, first + count - mid);
}
inline bool slab(const OBox& b, V3 O, V3 invD, float tmin, float tmax)
{
	float t0 = tmin, t1 = tmax;
	const float o[3] = { O.x, O.y, O.z }, id[3] = { invD.x, invD.y, invD.z };
	for (int k = 0; k < 3; ++k) {
		float a = (b.lo[k] - o[k]) * id[k], c = (b.hi[k] - o[k]) * id[k];
		if (a > c)
			std::swap(a, c);
		a = a - std::abs(a) * 4e-7f; // widen: never cull something the triangle test could accept
		c = c + std::abs(c) * 4e-7f;
		if (!(a != a))
			t0 = std::max(t0, a);
		if (!(c != c))
			t1 = std::min(t1, c);
	}
	return t0 <= t1;
}

void buildAccel(const Scene& sc, Accel& A)
{
	const prb_scene_desc& d = *sc.d;
	A.meshTris.resize(d.n_meshes);
	A.meshAccel.resize(d.n_meshes);
	for (uint32_t m = 0; m < d.n_meshes; ++m) {
		const prb_mesh& pm = d.meshes[m];
		auto& tris		   = A.meshTris[m];
		for (uint32_t f = 0; f < pm.face_count; ++f) {
			const uint32_t* idx = d.face_indices + 4 * (size_t)(pm.face_offset + f);
			const V3 v0 = ld3(d.vertices + 3 * (size_t)(pm.vertex_offset + idx[0])), v1 = ld3(d.vertices + 3 * (size_t)(pm.vertex_offset + idx[1])),
					 v2 = ld3(d.vertices + 3 * (size_t)(pm.vertex_offset + idx[2]));
			if (idx[3] != PRB_INVALID_ID) {
				const V3 v3 = ld3(d.vertices + 3 * (size_t)(pm.vertex_offset + idx[3]));
				faceTris(tris, v0, v1, v2, &v3, f);
			} else {
				faceTris(tris, v0, v1, v2, nullptr, f);
			}
		}
		if (tris.size() > 16) {
			OMeshAccel& acc = A.meshAccel[m];
			std::vector<OBox> boxes(tris.size());
			acc.tris.resize(tris.size());
			for (size_t i = 0; i < tris.size(); ++i) {
				boxes[i]	= triBox(tris[i]);
				acc.tris[i] = (uint32_t)i;
			}
			acc.nodes.reserve(tris.size());
			acc.nodes.push_back(ONode{});
			buildNode(acc, tris, boxes, 0, 0, (uint32_t)tris.size());
		}
	}
}

// closest hit over the whole scene; anyHit: stop at the first accepted hit
bool traceScene(const Scene& sc, const Accel& A, V3 O, V3 D, float tmin, float tmax, bool anyHit, Hit& best)
{
	const prb_scene_desc& d = *sc.d;
	bool found = false;
	auto consider = [&](uint32_t e, uint32_t prim, float t, float u, float v) {
		if (better(t, e, prim, best)) {
			best.entity = e;
			best.prim	= prim;
			best.t		= t;
			best.u		= u;
			best.v		= v;
		}
		found = true;
	};
	for (uint32_t e = 0; e < d.n_entities; ++e) {
		const prb_entity& en = d.entities[e];
		const float tfar	 = found ? best.t : tmax; // hits beyond the current best can never win
		if (en.type == PRB_ENTITY_SPHERE) {
			float t;
			if (sphereTest(O, D, tmin, tfar, ld3(en.geo), en.geo[3], t))
				consider(e, 0, t, 0, 0);
		} else if (en.type == PRB_ENTITY_PLANE) {
			const V3 v0 = ld3(en.geo + 14), v1 = ld3(en.geo + 17), v2 = ld3(en.geo + 20), v3 = ld3(en.geo + 23);
			float t, u, v;
			if (triTest(O, D, tmin, tfar, v0, v1, v3, t, u, v))
				consider(e, 0, t, u, v);
			if (triTest(O, D, tmin, found ? best.t : tmax, v2, v3, v1, t, u, v))
				consider(e, 0, t, 1 - u, 1 - v);
		} else { // mesh instance: ray into local space by the inverse transform, direction not re-normalised
			const V3 lo = xfPointE(en.world_to_local, O), ld = xfVecE(en.world_to_local, D);
			const auto& tris = A.meshTris[en.mesh_id];
			const OMeshAccel& acc = A.meshAccel[en.mesh_id];
			auto testTri = [&](const OTri& tr) {
				float t, u, v;
				if (triTest(lo, ld, tmin, found ? best.t : tmax, tr.a, tr.b, tr.c, t, u, v)) {
					if (tr.flags & 1) {
						u = 1 - u;
						v = 1 - v;
					}
					consider(e, tr.prim, t, u, v);
				}
			};
			if (acc.nodes.empty()) {
				for (const OTri& tr : tris) {
					testTri(tr);
					if (anyHit && found)
						return true;
				}
			} else {
				const V3 inv = mk(1.0f / ld.x, 1.0f / ld.y, 1.0f / ld.z);
				uint32_t stack[128];
				int sp		= 0;
				stack[sp++] = 0;
				while (sp > 0) {
					const ONode& n = acc.nodes[stack[--sp]];
					if (!slab(n.box, lo, inv, tmin, found ? best.t : tmax))
						continue;
					if (n.count > 0) {
						for (uint32_t i = n.first; i < n.first + n.count; ++i)
							testTri(tris[acc.tris[i]]);
						if (anyHit && found)
							return true;
					} else {
						stack[sp++] = n.left;
						stack[sp++] = n.right;
					}
				}
			}
		}
		if (anyHit && found)
			return true;
	}
	return found;
}

// ------------------------------------------------------------------ geometry point
struct GeomPoint { // GeometryPoint, src/core/geometry/GeometryPoint.h:10-25
	V3 N, Nx, Ny;
	float u, v;
	uint32_t entity, prim, material, emission;
};
struct FaceData {
	V3 V[4], N[4];
	float UV[4][2];
	bool quad;
	uint32_t slot;
};
FaceData getFace(const prb_scene_desc& d, const prb_mesh& m, uint32_t f)
{ // MeshBase::getFace, src/core/mesh/MeshBase.inl:96-134 (shared index set)
	FaceData fd{};
	const uint32_t* idx = d.face_indices + 4 * (size_t)(m.face_offset + f);
	fd.quad				= idx[3] != PRB_INVALID_ID;
	const int n			= fd.quad ? 4 : 3;
	for (int j = 0; j < n; ++j) {
		fd.V[j] = ld3(d.vertices + 3 * (size_t)(m.vertex_offset + idx[j]));
		if (m.features & PRB_MESH_HAS_NORMALS)
			fd.N[j] = ld3(d.normals + 3 * (size_t)(m.normal_offset + idx[j]));
		if (m.features & PRB_MESH_HAS_UVS) {
			fd.UV[j][0] = d.uvs[2 * (size_t)(m.uv_offset + idx[j])];
			fd.UV[j][1] = d.uvs[2 * (size_t)(m.uv_offset + idx[j]) + 1];
		}
	}
	fd.slot = d.face_slots[m.face_offset + f];
	return fd;
}
inline V3 triInterp(V3 v0, V3 v1, V3 v2, float u, float v) { return (v1 * u + v2 * v) + v0 * (1 - u - v); } // Triangle.h:23-27
inline V3 quadInterp(V3 v0, V3 v1, V3 v2, V3 v3, float u, float v)
{ // Quad::interpolate, src/core/geometry/Quad.h
	return ((v0 * (1 - u) * (1 - v) + v1 * u * (1 - v)) + v2 * (1 - u) * v) + v3 * u * v;
}
inline V3 faceInterpV(const FaceData& f, const V3* a, float u, float v) { return f.quad ? quadInterp(a[0], a[1], a[2], a[3], u, v) : triInterp(a[0], a[1], a[2], u, v); }
inline float faceArea(const FaceData& f)
{
	if (f.quad)
		return 0.5f * std::sqrt(norm2(cross(f.V[2] - f.V[0], f.V[3] - f.V[1])));
	return 0.5f * std::sqrt(norm2(cross(f.V[1] - f.V[0], f.V[2] - f.V[0])));
}

void provideGeometryPoint(const Scene& sc, uint32_t entityID, uint32_t prim, float qu, float qv, V3 position, GeomPoint& pt)
{
	const prb_scene_desc& d = *sc.d;
	const prb_entity& en	= d.entities[entityID];
	pt.entity				= entityID;
	pt.emission				= en.emission_id;
	if (en.type == PRB_ENTITY_MESH) { // mesh.cpp:205-250
		const prb_mesh& m = d.meshes[en.mesh_id];
		const FaceData f  = getFace(d, m, prim);
		if (m.features & PRB_MESH_HAS_NORMALS) {
			pt.N = faceInterpV(f, f.N, qu, qv);
			if (m.features & PRB_MESH_HAS_UVS) { // Face::tangentFromUV, Face.h:80-98
				const V3 dp1 = f.V[1] - f.V[0], dp2 = f.V[2] - f.V[0];
				const float du1 = f.UV[1][0] - f.UV[0][0], dv1 = f.UV[1][1] - f.UV[0][1];
				const float du2 = f.UV[2][0] - f.UV[0][0], dv2 = f.UV[2][1] - f.UV[0][1];
				const float det = diffProd(dv2, du1, dv1, du2);
				if (det <= PR_EPSILON) {
					tangent_frame(pt.N, pt.Nx, pt.Ny);
				} else {
					V3 nx = (dp1 * dv2 - dp2 * dv1) / det;
					nx	  = nx - pt.N * dot(pt.N, nx);
					nx	  = normalized(nx);
					pt.Nx = nx;
					pt.Ny = cross(pt.N, nx);
				}
			} else {
				frame_duff(pt.N, pt.Nx, pt.Ny);
			}
		} else { // rtcInterpolate dPdu, dPdv of the vertex buffer
			if (f.quad) {
				pt.Nx = (1 - qv) * (f.V[1] - f.V[0]) + qv * (f.V[2] - f.V[3]);
				pt.Ny = (1 - qu) * (f.V[3] - f.V[0]) + qu * (f.V[2] - f.V[1]);
			} else {
				pt.Nx = f.V[1] - f.V[0];
				pt.Ny = f.V[2] - f.V[0];
			}
			pt.N = cross(pt.Nx, pt.Ny);
		}
		if (m.features & PRB_MESH_HAS_UVS) {
			if (f.quad) {
				const float a = (1 - qu) * (1 - qv), b = qu * (1 - qv), c = (1 - qu) * qv, e = qu * qv;
				pt.u = ((f.UV[0][0] * a + f.UV[1][0] * b) + f.UV[2][0] * c) + f.UV[3][0] * e;
				pt.v = ((f.UV[0][1] * a + f.UV[1][1] * b) + f.UV[2][1] * c) + f.UV[3][1] * e;
			} else {
				pt.u = (f.UV[1][0] * qu + f.UV[2][0] * qv) + f.UV[0][0] * (1 - qu - qv);
				pt.v = (f.UV[1][1] * qu + f.UV[2][1] * qv) + f.UV[0][1] * (1 - qu - qv);
			}
		} else {
			pt.u = qu;
			pt.v = qv;
		}
		pt.material = f.slot < en.material_count ? d.entity_materials[en.material_offset + f.slot] : PRB_INVALID_ID;
		pt.N		= normalized(m3mul(en.normal_matrix, pt.N));
		pt.Nx		= normalized(m3mul(en.normal_matrix, pt.Nx));
		pt.Ny		= normalized(m3mul(en.normal_matrix, pt.Ny));
		pt.prim		= prim;
	} else if (en.type == PRB_ENTITY_SPHERE) { // sphere.cpp:128-143
		pt.N = normalized(position - xfPoint(en.local_to_world, mk(0, 0, 0)));
		tangent_frame(pt.N, pt.Nx, pt.Ny);
		uv_from_normal(pt.N, pt.u, pt.v);
		pt.prim		= 0;
		pt.material = d.entity_materials[en.material_offset];
	} else { // plane.cpp:206-220
		pt.N		= ld3(en.geo + 9);
		pt.Nx		= ld3(en.geo + 3);
		pt.Ny		= ld3(en.geo + 6);
		pt.u		= qu;
		pt.v		= qv;
		pt.prim		= 0;
		pt.material = d.entity_materials[en.material_offset];
	}
}

// ------------------------------------------------------------------ materials
constexpr float AIR = 1.0002926f; // dielectric.cpp:17
constexpr uint32_t MSF_Delta = 0x2, MSF_SpectralVarying = 0x4;
struct MatEval {
	Blob weight, pdf;
	uint32_t flags = 0, type = 0;
};
struct MatSample {
	V3 L;
	Blob weight, pdf;
	uint32_t flags = 0, type = 0;
	bool isDelta() const { return flags & MSF_Delta; }
	bool isHeroCollapsing() const { return (flags & MSF_Delta) && (flags & MSF_SpectralVarying); }
};
struct MatCtx { // MaterialSampleContext / MaterialEvalContext in shading space
	V3 V, L;
	Blob wvl;
	float u, v;
	uint32_t rayFlags;
};
inline uint32_t contribFlags(const prb_material& m) { return (m.flags & PRB_MATF_SPECTRAL_VARYING) ? MSF_SpectralVarying : 0; }
inline MatSample rejectSample(uint32_t type, uint32_t flags)
{
	MatSample s;
	s.L		 = mk(0, 0, 0);
	s.weight = blob(0);
	s.pdf	 = blob(0);
	s.type	 = type;
	s.flags	 = flags;
	return s;
}
inline RoughDistribution roughOf(const prb_material& m)
{
	RoughDistribution r;
	r.M1	= m.f[0];
	r.M2	= m.f[1];
	r.aniso = m.flags & PRB_MATF_ANISOTROPIC;
	r.vndf	= m.flags & PRB_MATF_VNDF;
	return r;
}

// --- principled closure, principled.cpp:34-447
struct Principled {
	Blob Base, IOR;
	float DiffuseTransmission, Roughness, Anisotropic, SpecularTransmission, SpecularTint, Flatness, Metallic, Sheen, SheenTint, Clearcoat, ClearcoatGloss;
	bool vndf, thin, hasTrans;
	const Scene* sc;
	static constexpr float EVAL_EPS = 1e-4f;
	static float mixf(float v0, float v1, float t) { return (1 - t) * v0 + t * v1; }
	static float schlickR0(float eta)
	{
		const float f = (eta - 1.0f) / (eta + 1.0f);
		return f * f;
	}
	float thinTransmissionRoughness() const { return std::max(0.0f, std::min(1.0f, (0.65f * (bsum(IOR) / 4) - 0.35f) * Roughness)); }
	RoughDistribution roughnessClosure(float r) const
	{
		const float aspect = std::sqrt(1 - Anisotropic * 0.9f);
		RoughDistribution d;
		d.M1	= std::max(0.001f, r * r / aspect);
		d.M2	= std::max(0.001f, r * r * aspect);
		d.aniso = true;
		d.vndf	= vndf;
		return d;
	}
	bool isDelta() const { return roughnessClosure(Roughness).isDelta(); }
	void lobes(V3 V, float& dr, float& dt, float& sr, float& st) const
	{ // calculateLobeDistribution :111-139
		dr = Roughness * Roughness * (1.0f - Metallic) * (1.0f - SpecularTransmission);
		sr = 1;
		if (hasTrans) {
			const float F = fresnel_dielectric(cosTheta(V), AIR, IOR[0]);
			dt			  = DiffuseTransmission * dr;
			st			  = (1.0f - F) * (1.0f - Metallic) * SpecularTransmission;
			sr *= F;
		} else {
			dt = 0;
			st = 0;
		}
		const float norm = dr + sr + dt + st;
		if (norm <= PR_EPSILON) {
			dr = 1;
			dt = sr = st = 0;
			return;
		}
		dr /= norm;
		sr /= norm;
		dt /= norm;
		st /= norm;
	}
	Blob tintColor(const Blob& wvl) const
	{ // :171-179
		float lum = 0;
		for (int i = 0; i < 4; ++i)
			lum = std::max(lum, Base[i] * cieEval(*sc, 1, wvl[i]));
		return lum > PR_EPSILON ? Base / lum : blob(1);
	}
	Blob disneyFresnelTerm(float HdotV, float HdotL, const Blob& wvl) const
	{ // :143-169
		Blob res;
		if (Metallic <= EVAL_EPS) {
			for (int i = 0; i < 4; ++i)
				res[i] = fresnel_dielectric(HdotV, AIR, IOR[i]);
			return res;
		}
		const Blob color = tintColor(wvl);
		for (int i = 0; i < 4; ++i) {
			const float eta = HdotV < 0 ? AIR / IOR[i] : IOR[i] / AIR;
			const float r0	= mixf(schlickR0(eta) * mixf(1.0f, color[i], SpecularTint), Base[i], Metallic);
			const float f1	= fresnel_dielectric(HdotV, AIR, IOR[i]);
			const float f2	= schlick(std::abs(HdotL), r0);
			res[i]			= mixf(f1, f2, Metallic);
		}
		return res;
	}
	float retroDiffuseTerm(const MatCtx& c, float HdotL) const
	{
		const float alpha2 = Roughness * Roughness;
		const float fd90   = 0.5f + 2 * HdotL * HdotL * alpha2;
		const float lk = schlick_term(absCosTheta(c.L)), vk = schlick_term(absCosTheta(c.V));
		return PR_INV_PI * fd90 * (lk + vk + lk * vk * (fd90 - 1.0f));
	}
	float subsurfaceTerm(const MatCtx& c, float HdotL) const
	{
		const float alpha2 = Roughness * Roughness;
		const float fss90  = HdotL * HdotL * alpha2;
		const float lk = schlick_term(absCosTheta(c.L)), vk = schlick_term(absCosTheta(c.V));
		const float fss = mixf(1.0f, fss90, lk) * mixf(1.0f, fss90, vk);
		const float f	= absCosTheta(c.L) + absCosTheta(c.V);
		if (std::abs(f) < PR_EPSILON)
			return 0.0f;
		return 1.25f * (fss * (1.0f / f - 0.5f) + 0.5f);
	}
	float diffuseTerm(const MatCtx& c, float HdotL) const
	{
		const float lk = schlick_term(absCosTheta(c.L)), vk = schlick_term(absCosTheta(c.V));
		float diffuse = 1;
		if (thin)
			diffuse = mixf(1.0f, subsurfaceTerm(c, HdotL), Flatness);
		return PR_INV_PI * diffuse * (1 - 0.5f * lk) * (1 - 0.5f * vk);
	}
	Blob specularReflectionTerm(const MatCtx& c, V3 H) const
	{
		MicrofacetReflection micro{ roughnessClosure(Roughness) };
		const float HdotV = dot(c.V, H), HdotL = dot(c.L, H);
		const Blob F = disneyFresnelTerm(HdotV, HdotL, c.wvl);
		return F * micro.eval(c.V, c.L);
	}
	float specularRefractionTermComponent(int i, const MatCtx& c) const
	{
		const float scaledR = thin ? thinTransmissionRoughness() : Roughness;
		MicrofacetTransmission micro{ roughnessClosure(scaledR), AIR, IOR[i] };
		const float R = micro.evalDielectric(c.V, c.L, false);
		return thin ? std::sqrt(Base[i]) * R : Base[i] * R;
	}
	float clearcoatTerm(const MatCtx& c, V3 H) const
	{
		const float F0 = 0.04f, R = 0.25f;
		const float D  = ndf_ggx(H, mixf(0.1f, 0.001f, ClearcoatGloss));
		const float hk = schlick_term(std::abs(dot(H, c.L)));
		const float F  = mixf(F0, 1.0f, hk);
		const float G  = g_1_smith_opt(absCosTheta(c.L), R) * g_1_smith_opt(absCosTheta(c.V), R);
		return R * D * F * G;
	}
	Blob sheenTerm(float HdotL, const Blob& wvl) const
	{
		if (Sheen <= EVAL_EPS)
			return blob(0);
		const Blob tint = tintColor(wvl);
		Blob sheenColor;
		for (int i = 0; i < 4; ++i)
			sheenColor[i] = mixf(1.0f, tint[i], SheenTint);
		return (sheenColor * Sheen) * schlick_term(std::abs(HdotL));
	}
	Blob eval(const MatCtx& c) const
	{ // :274-344
		if (absCosTheta(c.V) <= PR_EPSILON || absCosTheta(c.L) <= PR_EPSILON)
			return blob(0);
		const float diffuseWeight  = (1.0f - Metallic) * (1.0f - SpecularTransmission);
		const bool isTransmission  = !sameHemisphere(c.V, c.L);
		const bool upperHemisphere = cosTheta(c.V) >= 0.0f && !isTransmission;
		if (!hasTrans && isTransmission)
			return blob(0);
		const V3 rH		  = halfway_reflection(c.V, c.L);
		const float HdotL = dot(rH, c.L);
		Blob value		  = blob(0);
		if (diffuseWeight > EVAL_EPS) {
			if (!isTransmission) {
				const float retro = retroDiffuseTerm(c, HdotL) * diffuseWeight;
				const Blob sheen  = sheenTerm(HdotL, c.wvl) * diffuseWeight;
				value			  = value + (Base * retro + sheen) * absCosTheta(c.L);
			}
			if (!isTransmission) {
				const float diff = diffuseTerm(c, HdotL) * (thin ? 1 - DiffuseTransmission : diffuseWeight);
				value			 = value + Base * (diff * absCosTheta(c.L));
			}
			if (hasTrans && thin && isTransmission) {
				const float diff = diffuseTerm(c, HdotL) * DiffuseTransmission;
				value			 = value + Base * (diff * absCosTheta(c.L));
			}
		}
		value = value + specularReflectionTerm(c, rH);
		if (hasTrans) {
			const float transmissionWeight = (1.0f - Metallic) * SpecularTransmission;
			if (transmissionWeight > EVAL_EPS) {
				Blob weight;
				for (int i = 0; i < 4; ++i)
					weight[i] = specularRefractionTermComponent(i, c);
				if (c.rayFlags & PRB_RAY_LIGHT) {
					for (int i = 0; i < 4; ++i) {
						const float eta = HdotL < 0.0f ? IOR[i] / AIR : AIR / IOR[i];
						weight[i] *= eta * eta;
					}
				}
				value = value + weight * transmissionWeight;
			}
		}
		if (upperHemisphere && Clearcoat > EVAL_EPS)
			value = value + blob(clearcoatTerm(c, rH));
		return value;
	}
	Blob pdf(const MatCtx& c) const
	{ // :370-397
		if (absCosTheta(c.V) <= PR_EPSILON || absCosTheta(c.L) <= PR_EPSILON)
			return blob(0);
		float dr, dt, sr, st;
		lobes(c.V, dr, dt, sr, st);
		const bool isTransmission = !sameHemisphere(c.V, c.L);
		const float diffPdf		  = cos_hemi_pdf(absCosTheta(c.L));
		Blob pdfV				  = blob(0);
		if (!isTransmission) {
			pdfV = pdfV + blob(dr * diffPdf);
			if (sr > EVAL_EPS) {
				MicrofacetReflection refl{ roughnessClosure(Roughness) };
				pdfV = pdfV + blob(sr * refl.pdf(c.V, c.L));
			}
		}
		if (hasTrans && isTransmission) {
			pdfV = pdfV + blob(dt * diffPdf);
			if (st > EVAL_EPS) {
				const RoughDistribution rd = roughnessClosure(Roughness);
				Blob p;
				for (int i = 0; i < 4; ++i) {
					MicrofacetTransmission refr{ rd, AIR, IOR[i] };
					p[i] = refr.pdf(c.V, c.L);
				}
				pdfV = pdfV + p * st;
			}
		}
		return pdfV;
	}
	V3 sampleDiffuse(Rng& rnd, V3 V) const
	{
		const bool flip = cosTheta(V) < 0;
		const float u2	= rnd.getFloat(); // cos_hemi(getFloat(), getFloat()): right-to-left
		const float u1	= rnd.getFloat();
		const V3 L		= cos_hemi(u1, u2);
		return flip ? -L : L;
	}
	V3 sample(Rng& rnd, V3 V) const
	{ // :418-435
		if (absCosTheta(V) <= PR_EPSILON)
			return mk(0, 0, 0);
		float dr, dt, sr, st;
		lobes(V, dr, dt, sr, st);
		const float u0 = rnd.getFloat();
		if (u0 < dr)
			return sampleDiffuse(rnd, V);
		if (u0 < dr + dt)
			return -sampleDiffuse(rnd, V);
		float x, y;
		if (u0 < dr + dt + st) {
			MicrofacetTransmission refr{ roughnessClosure(Roughness), AIR, IOR[0] };
			rnd.get2D(x, y);
			return refr.sample(x, y, V);
		}
		MicrofacetReflection refl{ roughnessClosure(Roughness) };
		rnd.get2D(x, y);
		return refl.sample(x, y, V);
	}
};
Principled makePrincipled(const Scene& sc, const prb_material& m, const MatCtx& c)
{
	Principled p;
	p.sc	   = &sc;
	p.Base	   = evalNode(sc, m.node[0], c.wvl, c.u, c.v);
	p.IOR	   = evalNode(sc, m.node[1], c.wvl, c.u, c.v);
	p.vndf	   = m.flags & PRB_MATF_VNDF;
	p.thin	   = m.flags & PRB_MATF_THIN;
	p.hasTrans = m.flags & PRB_MATF_HAS_TRANSMISSION;
	p.DiffuseTransmission  = p.hasTrans ? m.f[PRB_PR_DIFF_TRANS] : 0.0f;
	p.SpecularTransmission = p.hasTrans ? m.f[PRB_PR_SPEC_TRANS] : 0.0f;
	p.Roughness			   = m.f[PRB_PR_ROUGHNESS];
	p.Anisotropic		   = m.f[PRB_PR_ANISOTROPIC];
	p.SpecularTint		   = m.f[PRB_PR_SPEC_TINT];
	p.Flatness			   = m.f[PRB_PR_FLATNESS];
	p.Metallic			   = m.f[PRB_PR_METALLIC];
	p.Sheen				   = m.f[PRB_PR_SHEEN];
	p.SheenTint			   = m.f[PRB_PR_SHEEN_TINT];
	p.Clearcoat			   = m.f[PRB_PR_CLEARCOAT];
	p.ClearcoatGloss	   = m.f[PRB_PR_CLEARCOAT_GLOSS];
	return p;
}

// --- rough dielectric closure, roughdielectric.cpp:42-137
struct RoughDielectric {
	RoughDistribution rd;
	Blob Spec, Trans, IOR;
	Blob eval(V3 V, V3 L, bool isLightPath) const
	{
		Blob w;
		if (sameHemisphere(V, L)) {
			MicrofacetReflection refl{ rd };
			for (int i = 0; i < 4; ++i)
				w[i] = refl.evalDielectric(V, L, AIR, IOR[i]);
			return w * Spec;
		}
		for (int i = 0; i < 4; ++i) {
			MicrofacetTransmission tr{ rd, AIR, IOR[i] };
			w[i] = tr.evalDielectric(V, L, isLightPath);
		}
		return w * Trans;
	}
	Blob pdf(V3 V, V3 L) const
	{
		Blob F, p;
		for (int i = 0; i < 4; ++i)
			F[i] = fresnel_dielectric(cosTheta(V), AIR, IOR[i]);
		if (sameHemisphere(V, L)) {
			MicrofacetReflection refl{ rd };
			for (int i = 0; i < 4; ++i)
				p[i] = refl.pdf(L, V);
			return F * p;
		}
		for (int i = 0; i < 4; ++i) {
			MicrofacetTransmission tr{ rd, AIR, IOR[i] };
			p[i] = tr.pdf(V, L);
		}
		Blob omf;
		for (int i = 0; i < 4; ++i)
			omf[i] = 1 - F[i];
		return omf * p;
	}
	V3 sample(Rng& rnd, V3 V) const
	{
		const float F = fresnel_dielectric(cosTheta(V), AIR, IOR[0]);
		float x, y;
		if (rnd.getFloat() <= F) {
			MicrofacetReflection refl{ rd };
			rnd.get2D(x, y);
			return refl.sample(x, y, V);
		}
		MicrofacetTransmission tr{ rd, AIR, IOR[0] };
		rnd.get2D(x, y);
		return tr.sample(x, y, V);
	}
};

// OrenNayarMaterial::calc (improved Oren-Nayar), orennayar.cpp:29-49
Blob orenNayarCalc(const Scene& sc, const prb_material& m, const MatCtx& c, V3 L, float NdotL)
{
	float roughness = m.f[0];
	roughness *= roughness;
	Blob weight = evalNode(sc, m.node[0], c.wvl, c.u, c.v);
	if (roughness > PR_EPSILON) {
		const float s = -NdotL * c.V.z + dot(c.V, L);
		const float t = s < PR_EPSILON ? 1.0f : std::max(NdotL, c.V.z);
		const Blob A  = blob(1 - 0.5f * roughness / (roughness + 0.33f)) + ((weight * 0.17f) * roughness) / (roughness + 0.13f);
		const float B = 0.45f * roughness / (roughness + 0.09f);
		weight		  = weight * (A + blob(B * s / t));
	}
	return weight;
}

void materialEvalLeaf(const Scene& sc, uint32_t matID, const MatCtx& c, MatEval& out)
{
	const prb_material& m = sc.d->materials[matID];
	out.flags			  = 0;
	switch (m.type) {
	case PRB_MAT_DIFFUSE: { // lambert.cpp:33-43
		const bool two	= m.flags & PRB_MATF_TWO_SIDED;
		const float d	= sameHemisphere(c.V, c.L) ? (two ? std::abs(c.L.z) : std::max(0.0f, c.L.z)) : 0;
		out.weight		= evalNode(sc, m.node[0], c.wvl, c.u, c.v) * d * PR_INV_PI;
		out.pdf			= blob(cos_hemi_pdf(d));
		out.type		= 0;
		break;
	}
	case PRB_MAT_DIELECTRIC: // dielectric.cpp:36-45 (never evaluated by the integrator: only-delta)
		out.pdf	   = blob(0);
		out.weight = blob(0);
		out.type   = 3;
		out.flags  = MSF_Delta | contribFlags(m);
		break;
	case PRB_MAT_MIRROR: // mirror.cpp:27-37 (never evaluated by the integrator: only-delta)
		out.pdf	   = blob(0);
		out.weight = blob(0);
		out.type   = 1;
		out.flags  = MSF_Delta;
		break;
	case PRB_MAT_ORENNAYAR: { // orennayar.cpp:51-60
		const float d = std::max(0.0f, c.L.z);
		out.weight	  = (orenNayarCalc(sc, m, c, c.L, d) * PR_INV_PI) * d;
		out.pdf		  = blob(cos_hemi_pdf(d));
		out.type	  = 0;
		break;
	}
	case PRB_MAT_CONDUCTOR: // conductor.cpp:33-42
		out.pdf	   = blob(0);
		out.weight = blob(0);
		out.type   = 1;
		out.flags  = MSF_Delta | contribFlags(m);
		break;
	case PRB_MAT_ROUGHCONDUCTOR: { // roughconductor.cpp:42-66
		out.type = 1;
		MicrofacetReflection closure{ roughOf(m) };
		if (closure.isDelta()) {
			out.pdf	   = blob(0);
			out.weight = blob(0);
			out.flags  = MSF_Delta | contribFlags(m);
			return;
		}
		const Blob eta = evalNode(sc, m.node[0], c.wvl, c.u, c.v), k = evalNode(sc, m.node[1], c.wvl, c.u, c.v);
		Blob factor;
		for (int i = 0; i < 4; ++i)
			factor[i] = closure.evalConductor(c.L, c.V, eta[i], k[i]);
		out.weight = evalNode(sc, m.node[2], c.wvl, c.u, c.v) * factor;
		out.pdf	   = blob(closure.pdf(c.L, c.V));
		out.flags  = contribFlags(m);
		break;
	}
	case PRB_MAT_ROUGHDIELECTRIC: { // roughdielectric.cpp:177-199
		RoughDielectric cl;
		cl.rd = roughOf(m);
		if (cl.rd.isDelta()) {
			out.pdf	   = blob(0);
			out.weight = blob(0);
			out.flags  = MSF_Delta | contribFlags(m);
			return;
		}
		cl.Spec	   = evalNode(sc, m.node[0], c.wvl, c.u, c.v);
		cl.Trans   = (m.flags & PRB_MATF_TRANSMISSION_COLOR) ? evalNode(sc, m.node[1], c.wvl, c.u, c.v) : cl.Spec;
		cl.IOR	   = evalNode(sc, m.node[2], c.wvl, c.u, c.v);
		out.weight = cl.eval(c.V, c.L, c.rayFlags & PRB_RAY_LIGHT);
		out.pdf	   = cl.pdf(c.V, c.L);
		out.type   = sameHemisphere(c.V, c.L) ? 1 : 3;
		out.flags  = contribFlags(m);
		break;
	}
	case PRB_MAT_PRINCIPLED: { // principled.cpp:496-523
		const Principled cl = makePrincipled(sc, m, c);
		if (cl.isDelta()) {
			out.weight = blob(0);
			out.pdf	   = blob(0);
			out.flags  = MSF_Delta;
			return;
		}
		if (sameHemisphere(c.V, c.L))
			out.type = cl.Roughness < 0.5f ? 1 : 0;
		else
			out.type = cl.Roughness < 0.5f ? 3 : 2;
		out.weight = cl.eval(c);
		out.pdf	   = cl.pdf(c);
		break;
	}
	}
}

void materialSampleLeaf(const Scene& sc, uint32_t matID, const MatCtx& c, Rng& rnd, MatSample& out)
{
	const prb_material& m = sc.d->materials[matID];
	out.flags			  = 0;
	switch (m.type) {
	case PRB_MAT_DIFFUSE: { // lambert.cpp:53-73
		if (!(m.flags & PRB_MATF_TWO_SIDED) && c.V.z < 0.0f) {
			out = rejectSample(0, 0);
			return;
		}
		const float u2 = rnd.getFloat(); // cos_hemi(RND.getFloat(), RND.getFloat()) evaluated right-to-left
		const float u1 = rnd.getFloat();
		out.L		   = cos_hemi(u1, u2);
		out.weight	   = evalNode(sc, m.node[0], c.wvl, c.u, c.v);
		out.pdf		   = blob(cos_hemi_pdf(out.L.z));
		out.type	   = 0;
		out.L		   = makeSameHemisphere(c.V, out.L);
		break;
	}
	case PRB_MAT_DIELECTRIC: { // dielectric.cpp:60-114
		out.pdf		  = blob(1);
		const Blob n2 = evalNode(sc, m.node[2], c.wvl, c.u, c.v);
		float F		  = fresnel_dielectric(cosTheta(c.V), AIR, n2[0]);
		const bool thin = m.flags & PRB_MATF_THIN;
		if (thin && F < 1.0f)
			F += (1 - F) * F / (F + 1);
		const Blob rWeight = evalNode(sc, m.node[0], c.wvl, c.u, c.v);
		if (rnd.getFloat() <= F) {
			out.type   = 1;
			out.L	   = reflect(c.V);
			out.weight = rWeight;
		} else {
			Blob tWeight = (m.flags & PRB_MATF_TRANSMISSION_COLOR) ? evalNode(sc, m.node[1], c.wvl, c.u, c.v) : rWeight;
			if (thin) {
				out.type   = 3;
				out.L	   = -c.V;
				out.weight = tWeight;
			} else {
				if (c.rayFlags & PRB_RAY_LIGHT) {
					const float eta = isPositiveHemisphere(c.V) ? AIR / n2[0] : n2[0] / AIR;
					tWeight			= tWeight * (eta * eta);
				}
				out.L = refract(AIR / n2[0], c.V);
				if (sameHemisphere(out.L, c.V)) {
					out.type   = 1;
					out.weight = rWeight;
				} else {
					out.type   = 3;
					out.weight = tWeight;
				}
			}
		}
		out.flags = MSF_Delta | contribFlags(m);
		break;
	}
	case PRB_MAT_MIRROR: // mirror.cpp:50-61
		out.weight = evalNode(sc, m.node[0], c.wvl, c.u, c.v);
		out.type   = 1;
		out.pdf	   = blob(1);
		out.L	   = reflect(c.V);
		out.flags  = MSF_Delta;
		break;
	case PRB_MAT_ORENNAYAR: { // orennayar.cpp:72-84
		const float u2 = rnd.getFloat(); // cos_hemi(RND.getFloat(), RND.getFloat()) evaluated right-to-left
		const float u1 = rnd.getFloat();
		out.L		   = cos_hemi(u1, u2);
		out.weight	   = orenNayarCalc(sc, m, c, out.L, std::max(0.0f, out.L.z));
		out.pdf		   = blob(cos_hemi_pdf(out.L.z));
		out.type	   = 0;
		break;
	}
	case PRB_MAT_CONDUCTOR: { // conductor.cpp:56-74
		const Blob eta = evalNode(sc, m.node[0], c.wvl, c.u, c.v), k = evalNode(sc, m.node[1], c.wvl, c.u, c.v);
		Blob fr;
		for (int i = 0; i < 4; ++i)
			fr[i] = fresnel_conductor(absCosTheta(c.V), 1, eta[i], k[i]);
		out.weight = fr * evalNode(sc, m.node[2], c.wvl, c.u, c.v);
		out.type   = 1;
		out.pdf	   = blob(1);
		out.L	   = reflect(c.V);
		out.flags  = MSF_Delta | contribFlags(m);
		break;
	}
	case PRB_MAT_ROUGHCONDUCTOR: { // roughconductor.cpp:82-117
		MicrofacetReflection closure{ roughOf(m) };
		float x, y;
		rnd.get2D(x, y);
		out.L	  = closure.sample(x, y, c.V);
		out.flags = contribFlags(m);
		if (closure.isDelta())
			out.flags |= MSF_Delta;
		if (!sameHemisphere(c.V, out.L)) {
			out = rejectSample(1, out.flags);
			return;
		}
		const Blob eta = evalNode(sc, m.node[0], c.wvl, c.u, c.v), k = evalNode(sc, m.node[1], c.wvl, c.u, c.v);
		Blob factor;
		for (int i = 0; i < 4; ++i)
			factor[i] = closure.evalConductor(out.L, c.V, eta[i], k[i]);
		out.weight = evalNode(sc, m.node[2], c.wvl, c.u, c.v) * factor;
		out.type   = 1;
		out.pdf	   = blob(closure.pdf(out.L, c.V));
		if (out.pdf[0] > PR_EPSILON)
			out.weight = out.weight / out.pdf[0];
		if (closure.isDelta())
			out.pdf = blob(1);
		break;
	}
	case PRB_MAT_ROUGHDIELECTRIC: { // roughdielectric.cpp:221-253
		RoughDielectric cl;
		cl.rd	  = roughOf(m);
		cl.Spec	  = evalNode(sc, m.node[0], c.wvl, c.u, c.v);
		cl.Trans  = (m.flags & PRB_MATF_TRANSMISSION_COLOR) ? evalNode(sc, m.node[1], c.wvl, c.u, c.v) : cl.Spec;
		cl.IOR	  = evalNode(sc, m.node[2], c.wvl, c.u, c.v);
		out.L	  = cl.sample(rnd, c.V);
		out.flags = contribFlags(m);
		if (cl.rd.isDelta())
			out.flags |= MSF_Delta;
		if (out.L.x == 0 && out.L.y == 0 && out.L.z == 0) {
			out = rejectSample(1, out.flags);
			return;
		}
		out.weight = cl.eval(c.V, out.L, c.rayFlags & PRB_RAY_LIGHT);
		out.pdf	   = cl.pdf(c.V, out.L);
		if (out.pdf[0] > PR_EPSILON)
			out.weight = out.weight / out.pdf[0];
		if (cl.rd.isDelta())
			out.pdf = blob(1);
		out.type = sameHemisphere(c.V, out.L) ? 1 : 3;
		break;
	}
	case PRB_MAT_PRINCIPLED: { // principled.cpp:536-590
		const Principled cl = makePrincipled(sc, m, c);
		out.L				= cl.sample(rnd, c.V);
		if (cl.isDelta())
			out.flags |= MSF_Delta;
		if (out.L.x == 0 && out.L.y == 0 && out.L.z == 0) {
			out = rejectSample(0, out.flags);
			return;
		}
		if (sameHemisphere(c.V, out.L))
			out.type = cl.Roughness < 0.5f ? 1 : 0;
		else
			out.type = cl.Roughness < 0.5f ? 3 : 2;
		MatCtx e   = c;
		e.L		   = out.L;
		out.weight = cl.eval(e);
		out.pdf	   = cl.pdf(e);
		if (out.pdf[0] > PR_EPSILON)
			out.weight = out.weight / out.pdf[0];
		if (cl.isDelta())
			out.pdf = blob(1);
		break;
	}
	}
}

// blend.cpp:20-148 / add.cpp:20-122 over two LEAF materials (node[0], node[1] hold their ids); everything else is a leaf
inline bool isCombination(const prb_material& m) { return m.type == PRB_MAT_BLEND || m.type == PRB_MAT_ADD; }
void materialEval(const Scene& sc, uint32_t matID, const MatCtx& c, MatEval& out)
{
	const prb_material& m = sc.d->materials[matID];
	if (!isCombination(m)) {
		materialEvalLeaf(sc, matID, c, out);
		return;
	}
	const bool add = m.type == PRB_MAT_ADD;
	const bool d0 = m.flags & PRB_MATF_CHILD0_DELTA, d1 = m.flags & PRB_MATF_CHILD1_DELTA;
	const float prob = std::min(1.0f, std::max(0.0f, m.f[0]));
	if (d0 && d1) { // MaterialDelta::All: never evaluated by the integrator
		out.pdf	   = blob(0);
		out.weight = blob(0);
		out.type   = 3;
		out.flags  = 0;
	} else if (d0 || d1) { // the non-delta child alone, scaled by its share (children may be combinations themselves)
		materialEval(sc, m.node[d0 ? 1 : 0], c, out);
		const float share = add ? 0.5f : (d0 ? prob : 1 - prob);
		out.pdf			  = out.pdf * share;
		if (!add)
			out.weight = out.weight * share;
	} else {
		MatEval o1, o2;
		materialEval(sc, m.node[0], c, o1);
		materialEval(sc, m.node[1], c, o2);
		out.flags = 0;
		if (add) {
			out.pdf	   = (o1.pdf + o2.pdf) / 2.0f;
			out.weight = o1.weight + o2.weight;
			out.type   = o1.type;
		} else {
			out.pdf	   = o1.pdf * (1 - prob) + o2.pdf * prob;
			out.weight = o1.weight * (1 - prob) + o2.weight * prob;
			out.type   = prob <= 0.5f ? o1.type : o2.type;
		}
	}
}
void materialSample(const Scene& sc, uint32_t matID, const MatCtx& c, Rng& rnd, MatSample& out)
{
	const prb_material& m = sc.d->materials[matID];
	if (!isCombination(m)) {
		materialSampleLeaf(sc, matID, c, rnd, out);
		return;
	}
	const bool add	 = m.type == PRB_MAT_ADD;
	const float prob = add ? 0.5f : std::min(1.0f, std::max(0.0f, m.f[0]));
	const bool first = rnd.getFloat() < (add ? 0.5f : 1 - prob);
	materialSample(sc, m.node[first ? 0 : 1], c, rnd, out);
	const float share = add ? 0.5f : (first ? 1 - prob : prob);
	if (!add)
		out.weight = out.weight * share;
	out.pdf = out.pdf * share;
}

// ------------------------------------------------------------------ samplers / mapper / camera
uint32_t mjPermute(uint32_t i, uint32_t l, uint32_t p)
{ // MultiJitteredSampler.cpp:21-76
	uint32_t w = l - 1;
	if (w == 0)
		return 0;
	const bool pow2 = (l & w) == 0;
	if (!pow2) {
		w |= w >> 1;
		w |= w >> 2;
		w |= w >> 4;
		w |= w >> 8;
		w |= w >> 16;
	}
	do {
		i ^= p;
		i *= 0xe170893d;
		i ^= p >> 16;
		i ^= (i & w) >> 4;
		i ^= p >> 8;
		i *= 0x0929eb3f;
		i ^= p >> 23;
		i ^= (i & w) >> 1;
		i *= 1 | p >> 27;
		i *= 0x6935fa69;
		i ^= (i & w) >> 11;
		i *= 0x74dcb303;
		i ^= (i & w) >> 2;
		i *= 0x9e501cc3;
		i ^= (i & w) >> 2;
		i *= 0xc860a3df;
		i &= w;
		i ^= i >> 5;
	} while (!pow2 && i >= l);
	return pow2 ? ((i + p) & w) : ((i + p) % l);
}
float haltonValue(uint32_t index, uint32_t base)
{ // HaltonSampler.cpp:14-25
	float result = 0;
	float f		 = 1;
	for (uint32_t i = index; i > 0;) {
		f = f / base;
		result += f * (i % base);
		i = static_cast<uint32_t>(std::floor(i / static_cast<float>(base)));
	}
	return result;
}
void sampler2D(const Scene& sc, const prb_sampler& s, Rng& rnd, uint32_t index, float& x, float& y)
{
	switch (s.type) {
	case PRB_SAMPLER_SOBOL: // SobolSampler.cpp:67-73
		if (s.max_samples <= index) {
			rnd.get2D(x, y);
		} else {
			const float* t = sc.d->pool + s.table_offset + s.max_samples;
			x			   = t[2 * index];
			y			   = t[2 * index + 1];
		}
		break;
	case PRB_SAMPLER_MJITT: { // MultiJitteredSampler.cpp:118-150 (PR_MJS_CLIP, PR_MJS_USE_RANDOM)
		constexpr uint32_t FH = 0x51633e2d, F1 = 0x68bc21eb, F2 = 0x02e5be93;
		const uint32_t maxS = std::max(1u, s.max_samples);
		index				= mjPermute(index, maxS, s.seed * FH);
		const uint32_t sx	= mjPermute(index % s.m2d_x, s.m2d_x, s.seed * F1);
		const uint32_t sy	= mjPermute(index / s.m2d_x, s.m2d_y, s.seed * F2);
		const float jx		= rnd.getFloat();
		const float jy		= rnd.getFloat();
		x					= (sx + (sy + jx) / s.m2d_y) / s.m2d_x;
		y					= (index + jy) / maxS;
		break;
	}
	case PRB_SAMPLER_HALTON: // HaltonSampler.cpp:52-60,100-108
		if (index < s.max_samples) {
			const float* t = sc.d->pool + s.table_offset + s.max_samples;
			x			   = t[2 * index];
			y			   = t[2 * index + 1];
		} else {
			x = haltonValue(index + s.seed, s.m2d_x);
			y = haltonValue(index + s.seed, s.m2d_y);
		}
		break;
	case PRB_SAMPLER_STRATIFIED: { // StratifiedSampler.cpp:29-36, Projection::stratified
		const float range = 1.0f / (int)s.m2d_x;
		const float ux	  = rnd.getFloat();
		x				  = ux * range + (int)(index % s.m2d_x) * range;
		const float uy	  = rnd.getFloat();
		y				  = uy * range + (int)(index / s.m2d_x) * range;
		break;
	}
	case PRB_SAMPLER_UNIFORM: x = y = 0.5f; break;
	default: rnd.get2D(x, y); break;
	}
}
float sampler1D(const Scene& sc, const prb_sampler& s, Rng& rnd, uint32_t index)
{
	switch (s.type) {
	case PRB_SAMPLER_SOBOL:
		if (s.max_samples <= index)
			return rnd.getFloat();
		return sc.d->pool[s.table_offset + index];
	case PRB_SAMPLER_MJITT: {
		const float j = rnd.getFloat();
		return (index % s.bins_1d + j) / s.bins_1d;
	}
	case PRB_SAMPLER_HALTON:
		return index < s.max_samples ? sc.d->pool[s.table_offset + index] : haltonValue(index + s.seed, s.m2d_x);
	case PRB_SAMPLER_STRATIFIED: { // StratifiedSampler.cpp:22-26
		const float range = 1.0f / (int)s.bins_1d;
		return rnd.getFloat() * range + (int)index * range;
	}
	case PRB_SAMPLER_UNIFORM: return 0.5f;
	default: return rnd.getFloat();
	}
}
// Distribution1D::sampleContinuous, src/base/math/Distribution1D.inl:76-86,119-135 (+ Interval::binary_search)
float sampleContinuous(const float* cdf, int size, float u, float& pdf)
{
	int first = 0, len = size;
	while (len > 0) {
		const int half = len / 2, middle = first + half;
		if (cdf[middle] <= u) {
			first = middle + 1;
			len -= half + 1;
		} else {
			len = half;
		}
	}
	const int off = std::max(0, std::min(first - 1, size - 2));
	float rem	  = u - cdf[off];
	const float k = cdf[off + 1] - cdf[off];
	if (k > PR_EPSILON)
		rem /= k;
	pdf = cdf[off + 1] - cdf[off];
	pdf *= (size - 1);
	return (off + rem) / (size - 1);
}
int sampleDiscrete(const float* cdf, int size, float u, float& pdf)
{
	int first = 0, len = size;
	while (len > 0) {
		const int half = len / 2, middle = first + half;
		if (cdf[middle] <= u) {
			first = middle + 1;
			len -= half + 1;
		} else {
			len = half;
		}
	}
	const int off = std::max(0, std::min(first - 1, size - 2));
	pdf			  = cdf[off + 1] - cdf[off];
	return off;
}
Blob heroWavelengths(float hero, float start, float end)
{ // constructHeroWavelength, src/plugins/main/spectralmapper/Standard.h:8-21
	const float span = end - start, delta = span / 4, s = hero - start;
	Blob w;
	w[0] = hero;
	for (int i = 1; i < 4; ++i)
		w[i] = start + std::fmod(s + i * delta, span);
	return w;
}

struct CameraSampleOut {
	V3 origin, dir;
	float tmin, tmax;
	Blob wvl, wvlPDF, importance;
	float blendWeight;
	bool mono;
};
void constructCameraRay(const Scene& sc, uint32_t px, uint32_t py, uint32_t iteration, Rng& rnd, CameraSampleOut& o)
{ // RenderTile::constructCameraRay, src/core/renderer/RenderTile.cpp:71-132
	const prb_scene_desc& d = *sc.d;
	const prb_settings& st	= d.settings;
	float ax, ay, lx, ly;
	sampler2D(sc, d.aa_sampler, rnd, iteration, ax, ay);
	const float pixx = ((float)px + ax) - 0.5f, pixy = ((float)py + ay) - 0.5f;
	sampler2D(sc, d.lens_sampler, rnd, iteration, lx, ly);
	const float time = st.time_alpha * sampler1D(sc, d.time_sampler, rnd, iteration) + st.time_beta;
	(void)time;
	o.blendWeight = 1.0f;
	if (st.spectral_mono) {
		o.wvl	 = blob(st.spectral_start);
		o.wvlPDF = blob(1.0f);
	} else {
		const float start = st.spectral_start, end = st.spectral_end;
		switch (d.pixel_mapper.type) {
		case PRB_MAPPER_SPD_CMIS: // spd.cpp:40-47
			for (int i = 0; i < 4; ++i) {
				float pdf;
				const float x = sampleContinuous(d.pool + d.pixel_mapper.cdf_offset, (int)d.pixel_mapper.cdf_size, rnd.getFloat(), pdf);
				o.wvl[i]	  = x * (end - start) + start;
				o.wvlPDF[i]	  = pdf;
			}
			break;
		case PRB_MAPPER_CIE: // cie.cpp:24-31,61-69 + CIE::sample_trunc, CIE.h:117-127 (0..1 for the full range)
			for (int i = 0; i < 4; ++i) {
				float pdf;
				const float cs = d.pixel_mapper.trunc_cdf_start, ce = d.pixel_mapper.trunc_cdf_end;
				const float v  = sampleContinuous(d.pool + d.pixel_mapper.cdf_offset, (int)d.pixel_mapper.cdf_size, cs + rnd.getFloat() * (ce - cs), pdf);
				pdf /= (ce - cs);
				o.wvl[i]	= v * (end - start) + start;
				o.wvlPDF[i] = pdf;
			}
			break;
		case PRB_MAPPER_AGH_CMIS: // agh.cpp:49-57: aghSample / aghPDF per wavelength
			for (int i = 0; i < 4; ++i) {
				const float C = d.pixel_mapper.trunc_cdf_start, N = d.pixel_mapper.trunc_cdf_end;
				o.wvl[i]	  = 538.0f - cr_atanh(C - N * rnd.getFloat()) / 0.0072f;
				const float K = cr_cosh(0.0072f * (o.wvl[i] - 538.0f));
				o.wvlPDF[i]	  = 1 / (K * K * N);
			}
			break;
		case PRB_MAPPER_AGH_HERO: { // agh.cpp:98-103 + Standard.h:8-21
			const float C = d.pixel_mapper.trunc_cdf_start, N = d.pixel_mapper.trunc_cdf_end;
			const float hero = 538.0f - cr_atanh(C - N * rnd.getFloat()) / 0.0072f;
			const float K	 = cr_cosh(0.0072f * (hero - 538.0f));
			const float span = end - start, delta = span / 4, s = hero - start;
			o.wvl[0] = hero;
			for (int i = 1; i < 4; ++i)
				o.wvl[i] = start + std::fmod(s + i * delta, span);
			o.wvlPDF = blob(1 / (K * K * N));
			break;
		}
		case PRB_MAPPER_SPD_HERO: { // spd.cpp:104-112
			float pdf;
			const float u	 = rnd.getFloat();
			const float hero = sampleContinuous(d.pool + d.pixel_mapper.cdf_offset, (int)d.pixel_mapper.cdf_size, u, pdf) * (end - start) + start;
			o.wvl			 = heroWavelengths(hero, start, end);
			o.wvlPDF		 = blob(pdf);
			break;
		}
		default: { // random.cpp:22-36
			const float u	 = rnd.getFloat();
			const float span = end - start, delta = span / 4, s = u * span;
			o.wvl[0] = s + start;
			for (int i = 1; i < 4; ++i)
				o.wvl[i] = start + std::fmod(s + i * delta, span);
			o.wvlPDF = blob(1.0f);
			break;
		}
		}
	}
	// PerspectiveCamera::constructRay, plugins/main/cameras/perspective.cpp:45-82
	const float nx = 2 * (pixx / (float)st.film_width - 0.5f);
	const float ny = -(2 * (pixy / (float)st.film_height - 0.5f));
	if (d.camera.type == PRB_CAMERA_ORTHOGRAPHIC) { // OrthoCamera::constructRay, plugins/main/cameras/ortho.cpp:47-66
		o.origin = (ld3(d.camera.origin) + ld3(d.camera.right) * nx) + ld3(d.camera.up) * ny;
		o.dir	 = ld3(d.camera.dir);
	} else {
		V3 dir	 = (ld3(d.camera.right) * nx + ld3(d.camera.up) * ny) + ld3(d.camera.dir);
		o.origin = ld3(d.camera.origin);
		if (d.camera.has_dof) { // PerspectiveCamera<HasDOF = true>::constructRay, perspective.cpp:66-75
			const float t = 2 * PR_PI * lx;
			const float s = cr_sin(t), c = cr_cos(t);
			const V3 e	  = (ld3(d.camera.aperture_x) * ly) * s + (ld3(d.camera.aperture_y) * ly) * c;
			o.origin	  = o.origin + e;
			dir			  = dir - e;
		}
		o.dir = normalized(dir);
	}
	o.tmin		   = d.camera.near_t;
	o.tmax		   = d.camera.far_t;
	o.importance   = blob(1.0f);
	o.mono		   = st.spectral_mono || !st.spectral_hero;
	if (o.mono)
		o.importance = o.importance * heroOnly();
}

// ------------------------------------------------------------------ the integrator
struct Film {
	std::vector<float> iterXYZ; // per-pixel sum of the current iteration (tile-local buffer of the reference)
	float* mean;				// running mean, W*H*3 (FrameOutputDevice::onEndOfIteration)
	uint32_t* sampleCount;
	float* aov; // 10 floats/pixel or null
	float* aovExt = nullptr; // PRB_AOV_EXT floats/pixel (tangent, bitangent, view, material id, emission id) or null
	uint32_t* feedback = nullptr; // AOV_Feedback, W*H words or null
	float* varMean	   = nullptr; // AOV_OnlineMean / AOV_OnlineVariance, W*H*3 each or null
	float* varVar	   = nullptr;
	// spectral channels restricted by a light path expression (prb_scene_desc::lpe): per expression a W*H*3 running mean at
	// lpeMean + k * W*H*3 and the sums of the running iteration, or null
	float* lpeMean = nullptr;
	std::vector<float> lpeIter;
};
struct Stats {
	uint64_t c[11] = {};
};
enum { S_CAMERA_RAY = 0, S_LIGHT_RAY, S_PRIMARY, S_BOUNCE, S_SHADOW, S_MONO, S_PIXEL_SAMPLE, S_ENTITY_HIT, S_BG_HIT, S_CAMERA_DEPTH, S_LIGHT_DEPTH };

struct PathState { // IntDirectInstance::TraversalContext, direct.cpp:47-57
	Blob Throughput = blob(1), PathPDF = blob(1), PrevPathPDF = blob(1), WavelengthPDF = blob(0);
	bool LastWasDelta = true, LastWasEmissive = false;
	V3 LastPosition = mk(0, 0, 0), LastNormal = mk(0, 0, 0);
};
struct RayS {
	V3 O, D;
	float tmin, tmax;
	Blob wvl;
	uint32_t depth, flags;
};
struct Group { // RayGroup, src/core/ray/RayGroup.h
	Blob importance, wvl, wvlPDF;
	float blendWeight;
};
struct IP { // IntersectionPoint, src/core/trace/IntersectionPoint.h:40-138
	V3 P, N, Nx, Ny;
	GeomPoint g;
	RayS ray;
	float NdotV, depth2;
};

inline float misTerm(bool power, float a) { return power ? a * a : a; } // vcm/MIS.h:7-30
inline Blob misTerm(bool power, Blob a) { return power ? a * a : a; }

// Test instrumentation: every fragment the integrator pushes, together with the path state and the pdfs its MIS weight was
// computed from, so that tests/test_mis_independent.py can recompute the weights from the text of direct.cpp alone.
enum { FK_BACKGROUND = 0, FK_DIRECT_HIT = 1, FK_NEE = 2, FK_INF_LIGHT = 3, FK_ZERO = 4 };
enum { FF_RAY_MONO = 1, FF_BSDF_MONO = 2, FF_LIGHT_DELTA = 4, FF_LAST_DELTA = 8, FF_LAST_EMISSIVE = 16, FF_FROM_BEHIND = 32, FF_VISIBLE = 64, FF_LIGHT_INFINITE = 128 };
struct FragLog { // 48 floats
	float kind, flags, depth, pixel;
	float mis[4], importance[4], radiance[4];
	float pathPDF[4], prevPathPDF[4], wvlPDF[4], bsdfPDF[4];
	float lightPdfS;   // NEE: the sampled light pdf (solid angle, x selection probability); direct hit: posPDF_S; inf light: sum_l pdf_S(l) is in extra
	float roulette;	   // NEE: mCameraRR.probability(pathLength)
	float extra;	   // inf light: number of non-delta infinite lights evaluated
	float accepted;	   // 1 when the film took the fragment, else -(feedback bits)
	float infPdfS[4];  // inf light: pdf_S (x selection probability) of up to four infinite lights
	float wvl[4];
	float groupImportance[4]; // RayGroup::Importance (HeroOnly for forced-monochrome renders)
};
static_assert(sizeof(FragLog) == 48 * sizeof(float), "FragLog layout is mirrored in tests/oracle_binding.py");

struct Integrator {
	const Scene& sc;
	const Accel& A;
	Film& film;
	Stats& stats;
	uint32_t pixelIndex;
	Group grp;
	std::vector<FragLog>* log = nullptr;
	FragLog pending{}; // filled by the handlers before pushSpectralFragment
	// LightPath mCameraPath (direct.cpp:67,472): the token string of the path so far, symbol = ScatteringType * 3 + ScatteringEvent
	// (LightPathToken.h:6-20)
	std::vector<uint8_t> path{};
	enum { TOK_CAMERA = 0 * 3 + 2, TOK_EMISSIVE = 1 * 3 + 2, TOK_BACKGROUND = 4 * 3 + 2 };
	static uint8_t scatterToken(uint32_t materialScatteringType)
	{ // LightPathToken(MaterialScatteringType), LightPathToken.h:40-60
		switch (materialScatteringType) {
		case 0: return 3 * 3 + 0; // DiffuseReflection  -> Reflection, Diffuse
		case 1: return 3 * 3 + 1; // SpecularReflection -> Reflection, Specular
		case 2: return 2 * 3 + 0; // DiffuseTransmission -> Refraction, Diffuse
		default: return 2 * 3 + 1; // SpecularTransmission -> Refraction, Specular
		}
	}
	// LightPathExpression::match (LPE_Automaton.h:17-33): the whole token string is walked from the start state
	bool lpeMatches(uint32_t k) const
	{
		const prb_scene_desc& d = *sc.d;
		const prb_lpe& l		= d.lpe[k];
		uint32_t state			= l.start_state;
		for (uint8_t sym : path) {
			state = d.lpe_tables[l.next_offset + state * PRB_LPE_SYMBOLS + sym];
			if (state == PRB_LPE_REJECT)
				return false;
		}
		return d.lpe_tables[l.final_offset + state] != 0;
	}

	// LocalFrameOutputDevice::commitSpectrals2, src/loader/output/LocalFrameOutputDevice.cpp:88-164 (filter applied later)
	void pushSpectralFragment(const Blob& mis, const Blob& importance, const Blob& radiance, uint32_t rayFlags)
	{
		// a monotonic film (FrameOutputDevice(..., spectralMono), loader/Environment.cpp:194-198) instantiates
		// commitSpectrals2<IsMono = true>: every fragment is hero-only and mapSpectral<true> (:76-85) stores the unweighted
		// hero sample in all three channels
		const bool monotonic  = sc.d->settings.film_monotonic;
		const bool isMono	  = monotonic || (rayFlags & PRB_RAY_MONOCHROME);
		const Blob heroFactor = isMono ? heroOnly() : blob(1);
		const Blob imp		  = grp.importance * importance;
		const Blob contrib	  = heroFactor * ((mis * imp) * radiance);
		uint32_t feedback	  = 0; // LocalFrameOutputDevice.cpp:125-143
		for (int i = 0; i < 4; ++i) {
			if (std::isnan(contrib[i]))
				feedback |= PRB_FEEDBACK_NAN;
			if (std::isinf(contrib[i]))
				feedback |= PRB_FEEDBACK_INFINITE;
			if (contrib[i] < -PR_EPSILON)
				feedback |= PRB_FEEDBACK_NEGATIVE;
		}
		if (log) {
			FragLog e  = pending;
			e.pixel	   = (float)pixelIndex;
			e.accepted = feedback ? -(float)feedback : 1.0f;
			for (int i = 0; i < 4; ++i) {
				e.mis[i]		= mis[i];
				e.importance[i] = imp[i];
				e.radiance[i]	= radiance[i];
				e.wvl[i]		= grp.wvl[i];
				e.groupImportance[i] = grp.importance[i];
			}
			if (rayFlags & PRB_RAY_MONOCHROME)
				e.flags = (float)((uint32_t)e.flags | FF_RAY_MONO);
			log->push_back(e);
			pending = FragLog{};
		}
		if (feedback) {
			if (film.feedback)
				film.feedback[pixelIndex] |= feedback;
			return;
		}
		float xyz[3] = { 0, 0, 0 };
		if (monotonic) {
			xyz[0] = xyz[1] = xyz[2] = contrib[0];
		} else {
			for (int k = 0; k < 4; ++k)
				for (int c = 0; c < 3; ++c)
					xyz[c] += contrib[k] * cieEval(sc, c, grp.wvl[k]);
		}
		for (int c = 0; c < 3; ++c)
			film.iterXYZ[3 * (size_t)pixelIndex + c] += grp.blendWeight * xyz[c];
		if (film.lpeMean) { // LocalFrameOutputDevice.cpp:100-111: every expression that matches the fragment's path takes the triplet too
			const size_t stride = (size_t)sc.d->settings.film_width * sc.d->settings.film_height * 3;
			for (uint32_t k = 0; k < sc.d->n_lpe; ++k)
				if (lpeMatches(k))
					for (int c = 0; c < 3; ++c)
						film.lpeIter[k * stride + 3 * (size_t)pixelIndex + c] += grp.blendWeight * xyz[c];
		}
	}
	void logState(int kind, uint32_t flags, uint32_t depth, const PathState& cur)
	{
		if (!log)
			return;
		pending		  = FragLog{};
		pending.kind  = (float)kind;
		pending.flags = (float)(flags | (cur.LastWasDelta ? FF_LAST_DELTA : 0) | (cur.LastWasEmissive ? FF_LAST_EMISSIVE : 0));
		pending.depth = (float)depth;
		for (int i = 0; i < 4; ++i) {
			pending.pathPDF[i]	   = cur.PathPDF[i];
			pending.prevPathPDF[i] = cur.PrevPathPDF[i];
			pending.wvlPDF[i]	   = cur.WavelengthPDF[i];
		}
	}
	float rrProbability(uint32_t pathLength, bool delta) const
	{ // RussianRoulette::probability, vcm/RussianRoulette.h:22-34
		if (pathLength == 0 || delta)
			return 1.0f;
		return sc.rrProb[std::min<size_t>(pathLength, sc.rrProb.size() - 1)];
	}

	void handleDirectHit(const IP& ip, PathState& cur)
	{ // direct.cpp:355-412
		const prb_scene_desc& d = *sc.d;
		if (ip.g.emission >= d.n_emissions)
			return;
		const float cosC = -ip.NdotV;
		if (std::abs(cosC) <= PR_EPSILON)
			return;
		const bool hitFromBehind = cosC < 0.0f;
		const Blob radiance		 = hitFromBehind ? blob(0) : evalNode(sc, d.emissions[ip.g.emission].radiance_node, ip.ray.wvl, ip.g.u, ip.g.v);
		const bool mono			 = ip.ray.flags & PRB_RAY_MONOCHROME;
		const Blob heroFactor	 = mono ? heroOnly() : blob(1);
		logState(FK_DIRECT_HIT, hitFromBehind ? FF_FROM_BEHIND : 0, ip.ray.depth, cur);
		if (!d.settings.do_nee || hitFromBehind || cur.LastWasDelta) {
			path.push_back(TOK_EMISSIVE); // direct.cpp:387-389
			pushSpectralFragment(heroFactor / (cur.WavelengthPDF * bsum(heroFactor)), cur.Throughput, radiance, ip.ray.flags);
			path.pop_back();
			return;
		}
		const prb_entity& en = d.entities[ip.g.entity];
		const float selProb	 = en.light_id != PRB_INVALID_ID ? d.lights[en.light_id].select_pdf : 0.0f;
		float posPDF		 = 0;
		bool isArea			 = false;
		if (en.light_id != PRB_INVALID_ID) {
			isArea = true;
			posPDF = entityPositionPDF(ip.g.entity, ip.P, cur.LastPosition);
		}
		if (isArea)
			posPDF = posPDF * ip.depth2 / std::abs(cosC); // IS::toSolidAngle
		const float posPDF_S = posPDF * selProb;
		const bool power	 = d.settings.mis_power;
		const float denom	 = bsum(misTerm(power, cur.PrevPathPDF * posPDF_S)) + bsum(misTerm(power, cur.PathPDF));
		const Blob mis		 = (heroFactor * misTerm(power, cur.PathPDF[0])) / (misTerm(power, cur.WavelengthPDF) * denom);
		pending.lightPdfS	 = posPDF_S;
		path.push_back(TOK_EMISSIVE); // direct.cpp:409-411
		pushSpectralFragment(mis, cur.Throughput, radiance, ip.ray.flags);
		path.pop_back();
	}

	// IEntity::sampleParameterPointPDF(p, info): mesh default 1/worldArea; sphere 2*pdfCache; plane spherical rectangle
	float entityPositionPDF(uint32_t entityID, V3 p, V3 infoOrigin) const
	{
		const prb_entity& en = sc.d->entities[entityID];
		if (en.type == PRB_ENTITY_SPHERE)
			return 2 * en.geo[5];
		if (en.type == PRB_ENTITY_PLANE) { // plane.cpp:184-195
			const SQ sq		  = computeSQ(en, infoOrigin);
			const float pdf_s = sq.S > PR_EPSILON ? 1 / sq.S : 0.0f;
			const V3 L		  = p - infoOrigin;
			const float dist2 = norm2(L);
			const float ndotv = std::abs(dot(normalized(L), ld3(en.geo + 26)));
			return ndotv <= PR_EPSILON ? 0 : pdf_s * std::abs(ndotv) / dist2;
		}
		return en.pdf_area;
	}
	struct SQ { // plane.cpp:100-145 (Urena et al. spherical rectangle)
		V3 o, n;
		float z0, x0, y0, x1, y1, b0, b1, k, S;
	};
	static float safe_acos(float a) { return cr_acos(std::max(-1.0f, std::min(1.0f, a))); }
	SQ computeSQ(const prb_entity& en, V3 o) const
	{
		SQ sq;
		const V3 mS = ld3(en.geo), mEx = ld3(en.geo + 3), mEy = ld3(en.geo + 6), mEz = ld3(en.geo + 9);
		sq.o	   = o;
		sq.n	   = mEz;
		const V3 d = mS - sq.o;
		sq.x0	   = dot(d, mEx);
		sq.y0	   = dot(d, mEy);
		sq.z0	   = dot(d, sq.n);
		sq.x1	   = sq.x0 + en.geo[12];
		sq.y1	   = sq.y0 + en.geo[13];
		if (sq.z0 > 0.0f) {
			sq.z0 = -sq.z0;
			sq.n  = -sq.n;
		}
		const float a[4] = { sq.x0, sq.y1, sq.x1, sq.y0 }, b[4] = { sq.x1, sq.y0, sq.x0, sq.y1 }, c[4] = { sq.y0, sq.x1, sq.y1, sq.x0 };
		float nz[4];
		for (int i = 0; i < 4; ++i) {
			const float diff = a[i] - b[i];
			nz[i]			 = c[i] * diff;
			nz[i] /= std::sqrt(sq.z0 * sq.z0 * diff * diff + nz[i] * nz[i]);
		}
		const float g0 = safe_acos(-nz[0] * nz[1]), g1 = safe_acos(-nz[1] * nz[2]), g2 = safe_acos(-nz[2] * nz[3]), g3 = safe_acos(-nz[3] * nz[0]);
		sq.b0 = nz[0];
		sq.b1 = nz[2];
		sq.k  = 2 * PR_PI - g2 - g3;
		sq.S  = g0 + g1 - sq.k;
		return sq;
	}

	struct LightSample {
		Blob radiance;
		V3 outgoing, lightPos;
		float posPDF, dirPDF_S, cosLight;
		bool posIsArea, infinite, delta;
	};
	// ---- sky / sun (plugins/main/infinitelights/sky.cpp, sun.cpp; tables precomputed by the host, see prb200_abi.h)
	static constexpr float SKY_ELEVATION_RANGE = PR_PI * 0.5f; // skysun/ElevationAzimuth.h:6-7
	static constexpr float SKY_AZIMUTH_RANGE   = PR_PI * 2;
	float skyModelRadiance(const prb_light& l, int band, float el, float az) const
	{ // SkyModel::radiance, skysun/SkyModel.h:19-24
		const int azc = (int)l.az_count, elc = (int)l.el_count;
		const int az_in = std::max(0, std::min<int>(azc - 1, int(az / SKY_AZIMUTH_RANGE * (float)azc)));
		const int el_in = std::max(0, std::min<int>(elc - 1, int(el / SKY_ELEVATION_RANGE * (float)elc)));
		return sc.d->pool[l.table_offset + ((size_t)el_in * azc + az_in) * PRB_SKY_BANDS + band];
	}
	Blob skyRadiance(const prb_light& l, const Blob& wvls, float el, float az) const
	{ // SkyLight::radiance, sky.cpp:168-184
		Blob b;
		for (int i = 0; i < 4; ++i) {
			const float af	= std::max(0.0f, (wvls[i] - PRB_SKY_BAND_START) / PRB_SKY_BAND_DELTA);
			const int index = (int)std::min<float>(PRB_SKY_BANDS - 2, af);
			const float t	= std::min<float>(PRB_SKY_BANDS - 1, af) - index;
			b[i]			= skyModelRadiance(l, index, el, az) * (1 - t) + skyModelRadiance(l, index + 1, el, az) * t;
		}
		return b;
	}
	// Distribution2D::sampleContinuous / continuousPdf, core/sampler/Distribution2D.cpp:13-31
	void dist2DSampleContinuous(const prb_light& l, float u0, float u1, float& d0, float& d1, float& pdf) const
	{
		const float* marginal = sc.d->pool + l.dist_offset;
		const int h = (int)l.dist_h, w = (int)l.dist_w;
		float pdf1, pdf0;
		d1 = sampleContinuous(marginal, h + 1, u1, pdf1);
		// offset found by the marginal's sampleDiscrete
		int first = 0, len = h + 1;
		while (len > 0) {
			const int half = len / 2, middle = first + half;
			if (marginal[middle] <= u1) {
				first = middle + 1;
				len -= half + 1;
			} else {
				len = half;
			}
		}
		const int moff		   = std::max(0, std::min(first - 1, h - 1));
		const float* conditional = marginal + (h + 1) + (size_t)moff * (w + 1);
		d0					   = sampleContinuous(conditional, w + 1, u0, pdf0);
		pdf					   = pdf0 * pdf1;
	}
	float dist2DContinuousPdf(const prb_light& l, float x0, float x1) const
	{ // Distribution1D::continuousPdf, Distribution1D.inl:93-99
		const float* marginal = sc.d->pool + l.dist_offset;
		const size_t h = l.dist_h, w = l.dist_w;
		const size_t moff		 = std::min<size_t>(h - 1, (size_t)(x1 * (float)h));
		const float pdf1		 = (marginal[moff + 1] - marginal[moff]) * (float)h;
		const float* conditional = marginal + (h + 1) + moff * (w + 1);
		const size_t off		 = std::min<size_t>(w - 1, (size_t)(x0 * (float)w));
		const float pdf0		 = (conditional[off + 1] - conditional[off]) * (float)w;
		return pdf0 * pdf1;
	}
	// IInfiniteLight::eval for every infinite light type
	void infLightEval(const prb_light& l, const RayS& ray, Blob& rad, float& pdfS) const
	{
		if (l.type == PRB_LIGHT_SKY) { // SkyLight::eval, sky.cpp:53-80
			float theta, phi;
			spherical_from_direction(m3mul(l.inv_normal_matrix, ray.D), theta, phi);
			const float el = 0.5f * PR_PI - theta; // ElevationAzimuth::fromThetaPhi (phi is already in [0, 2 pi))
			float az	   = phi;
			if (az < 0)
				az += 2 * PR_PI;
			if (!l.sky_extend && el < 0) {
				rad	 = blob(0);
				pdfS = 0;
				return;
			}
			rad = skyRadiance(l, ray.wvl, el, az);
			pdfS = l.sky_extend ? dist2DContinuousPdf(l, az / SKY_AZIMUTH_RANGE, el / (2 * SKY_ELEVATION_RANGE) + 0.5f)
								: dist2DContinuousPdf(l, az / SKY_AZIMUTH_RANGE, el / SKY_ELEVATION_RANGE);
			const float f	  = cr_cos(el);
			const float denom = 2 * PR_PI * PR_PI * f;
			pdfS *= (denom <= PR_EPSILON) ? 0.0f : 1.0f / denom;
		} else if (l.type == PRB_LIGHT_SUN) { // SunLight::eval, sun.cpp:60-75
			const float cosine = std::max(0.0f, dot(ray.D, ld3(l.sun_dir)));
			if (cosine < l.sun_cos_theta) {
				rad	 = blob(0);
				pdfS = 0;
			} else {
				for (int i = 0; i < 4; ++i)
					rad[i] = tableLookup(sc.d->pool + l.table_offset, l.table_count, l.table_start, l.table_end, ray.wvl[i]);
				pdfS = l.sun_pdf;
			}
		} else {
			envEval(l, ray, rad, pdfS);
		}
	}
	static bool isInfLight(const prb_light& l) { return l.type != PRB_LIGHT_AREA; }
	static bool isDeltaLight(const prb_light& l) { return l.type == PRB_LIGHT_SUN_DELTA; }

	// Light::sample with SamplingInfo + Point (NEE), src/core/light/Light.cpp:108-226
	void sampleLight(const prb_light& l, const IP& ip, Rng& rnd, LightSample& o)
	{
		const prb_scene_desc& d = *sc.d;
		o.delta = false;
		if (l.type == PRB_LIGHT_SKY) { // SkyLight::sampleDir / samplePosDir, sky.cpp:82-113
			float dx, dy, px, py;
			rnd.get2D(dx, dy);
			rnd.get2D(px, py);
			float pdf;
			float u0, u1;
			dist2DSampleContinuous(l, dx, dy, u0, u1, pdf);
			float el, az;
			if (l.sky_extend) {
				el = 2 * SKY_ELEVATION_RANGE * (u1 - 0.5f);
				az = SKY_AZIMUTH_RANGE * u0;
			} else {
				el = SKY_ELEVATION_RANGE * u1;
				az = SKY_AZIMUTH_RANGE * u0;
			}
			const float theta = 0.5f * PR_PI - el; // ElevationAzimuth::toDirection
			o.outgoing		  = m3mul(l.normal_matrix, spherical_cartesian(cr_sin(theta), cr_cos(theta), cr_sin(az), cr_cos(az)));
			const float f	  = cr_cos(el);
			const float denom = 2 * PR_PI * PR_PI * f;
			o.dirPDF_S		  = pdf * ((denom <= PR_EPSILON) ? 0.0f : 1.0f / denom);
			o.radiance		  = skyRadiance(l, ip.ray.wvl, el, az);
			o.lightPos		  = ip.P + l.scene_radius * o.outgoing;
			o.posPDF		  = 1;
			o.posIsArea		  = true;
			o.cosLight		  = 1;
			o.infinite		  = true;
			return;
		}
		if (l.type == PRB_LIGHT_SUN || l.type == PRB_LIGHT_SUN_DELTA) { // SunLight / SunDeltaLight::sampleDir, sun.cpp:77-99,177-199
			float dx, dy, px, py;
			rnd.get2D(dx, dy);
			rnd.get2D(px, py);
			const V3 sunDir = ld3(l.sun_dir);
			if (l.type == PRB_LIGHT_SUN) {
				// Sampling::uniform_cone, src/base/math/Sampling.h:101-107
				const float cosTheta = std::fma(dx, l.sun_cos_theta, 1 - dx);
				const float sinTheta = std::sqrt(std::max(0.0f, diffProd(1, 1, cosTheta, cosTheta)));
				const float phi		 = 2 * PR_PI * dy;
				const V3 local		 = mk(cr_cos(phi) * sinTheta, cr_sin(phi) * sinTheta, cosTheta);
				o.outgoing			 = fromTangentSpace(sunDir, ld3(l.sun_dx), ld3(l.sun_dy), local);
				o.dirPDF_S			 = l.sun_pdf;
			} else {
				o.outgoing = sunDir;
				o.dirPDF_S = 1;
				o.delta	   = true;
			}
			for (int i = 0; i < 4; ++i)
				o.radiance[i] = tableLookup(sc.d->pool + l.table_offset, l.table_count, l.table_start, l.table_end, ip.ray.wvl[i]);
			o.lightPos	= ip.P + l.scene_radius * o.outgoing;
			o.posPDF	= 0; // Position_PDF_A is left at its default (0) when a point is given, InfiniteLightSamplePosDirOutput
			o.posIsArea = true;
			o.cosLight	= 1;
			o.infinite	= true;
			return;
		}
		if (l.type == PRB_LIGHT_ENV) { // environment.cpp sampleDir/samplePosDir, non-distribution branch
			float dx, dy, px, py;
			rnd.get2D(dx, dy);
			rnd.get2D(px, py);
			V3 local;
			if (l.dist_w) { // EnvironmentLight<UseDistribution = true>::sampleDir, environment.cpp:83-104
				float u0, u1, pdf;
				dist2DSampleContinuous(l, dx, dy, u0, u1, pdf);
				local				 = cartesian_from_uv(u0, u1);
				const float sinTheta = cr_sin(u1 * PR_PI);
				const float denom	 = 2 * PR_PI * PR_PI * sinTheta;
				o.dirPDF_S			 = pdf * ((denom <= PR_EPSILON) ? 0.0f : 1.0f / denom);
				dx					 = u0; // coord.UV = uv
				dy					 = u1;
			} else {
				local	   = cos_hemi(dx, dy);
				o.dirPDF_S = cos_hemi_pdf(local.z);
			}
			o.outgoing	   = m3mul(l.normal_matrix, local);
			o.radiance	   = evalNode(sc, l.radiance_node, ip.ray.wvl, dx, dy);
			o.lightPos	   = ip.P + l.scene_radius * o.outgoing;
			o.posPDF	   = 1;
			o.posIsArea	   = true;
			o.cosLight	   = 1;
			o.infinite	   = true;
			return;
		}
		o.infinite			 = false;
		const prb_entity& en = d.entities[l.entity_id];
		float rx, ry;
		rnd.get2D(rx, ry);
		V3 pos;
		float su, sv, pdfA;
		uint32_t prim = 0;
		if (en.type == PRB_ENTITY_MESH) { // mesh.cpp:187-203
			const prb_mesh& m = d.meshes[en.mesh_id];
			float k1, k2;
			const float f1	= std::modf(rx * m.face_count, &k1); // SplitSample1D, src/base/math/SplitSample.h:6-26
			const float f2	= std::modf(ry * m.face_count, &k2);
			const uint32_t faceID = std::min<uint32_t>((uint32_t)k1, m.face_count - 1);
			const FaceData f	  = getFace(d, m, faceID);
			pdfA				  = 1.0f / (m.face_count * faceArea(f) * en.jacobian_det);
			if (!f.quad) { // Triangle::sample, Triangle.h:46-55
				if (f2 > f1) {
					const float x = f1 / 2;
					su			  = x;
					sv			  = f2 - x;
				} else {
					const float y = f2 / 2;
					su			  = f1 - y;
					sv			  = y;
				}
			} else {
				su = f1;
				sv = f2;
			}
			pos	 = xfPoint(en.local_to_world, faceInterpV(f, f.V, su, sv));
			prim = faceID;
		} else if (en.type == PRB_ENTITY_SPHERE) { // sphere.cpp:106-116
			V3 n			 = cartesian_from_uv(rx, ry);
			const V3 local_o = normalized(xfPoint(en.world_to_local, ip.P));
			if (dot(local_o, n) < -PR_EPSILON)
				n = -n;
			pos = xfPoint(en.local_to_world, en.geo[4] * n);
			uv_from_normal(n, su, sv);
			pdfA = 2 * en.geo[5];
		} else { // plane.cpp:147-182
			const SQ sq	   = computeSQ(en, ip.P);
			const V3 mEx = ld3(en.geo + 3), mEy = ld3(en.geo + 6);
			const float au = std::fma(rx, sq.S, sq.k);
			const float fu = std::fma(cr_cos(au), sq.b0, -sq.b1) / cr_sin(au);
			const float cu = std::min(1.0f, std::max(-1.0f, std::copysign(1.0f, fu) / std::sqrt(sumProd(fu, fu, sq.b0, sq.b0))));
			const float xu = std::min(sq.x1, std::max(sq.x0, -(cu * sq.z0) / std::max(1e-7f, std::sqrt(std::fma(-cu, cu, 1.0f)))));
			const float dd = std::sqrt(sumProd(xu, xu, sq.z0, sq.z0));
			const float h0 = sq.y0 / std::sqrt(sumProd(dd, dd, sq.y0, sq.y0));
			const float h1 = sq.y1 / std::sqrt(sumProd(dd, dd, sq.y1, sq.y1));
			const float hv = std::fma(ry, h1 - h0, h0);
			const float hv2 = hv * hv;
			const float yv	= (hv2 < 1.0f - 1e-6f) ? (hv * dd) / std::sqrt(1.0f - hv2) : sq.y1;
			pos				= ((sq.o + xu * mEx) + yv * mEy) + sq.z0 * sq.n;
			const float pdf_s = sq.S > PR_EPSILON ? 1 / sq.S : 0.0f;
			const V3 L		  = pos - ip.P;
			const float dist2 = norm2(L);
			const float ndotv = std::abs(dot(normalized(L), ld3(en.geo + 26)));
			pdfA			  = ndotv <= PR_EPSILON ? 0 : pdf_s * ndotv / dist2; // IS::toArea
			const V3 lp		  = xfPoint(en.world_to_local, pos) - ld3(en.geo + 29); // Plane::project
			su				  = dot(ld3(en.geo + 32), lp) * en.geo[38];
			sv				  = dot(ld3(en.geo + 35), lp) * en.geo[39];
		}
		GeomPoint gp;
		provideGeometryPoint(sc, l.entity_id, prim, su, sv, pos, gp);
		o.outgoing	= normalized(pos - ip.P);
		o.dirPDF_S	= 1;
		o.cosLight	= std::min(1.0f, std::max(-1.0f, -dot(o.outgoing, gp.N)));
		o.radiance	= evalNode(sc, d.emissions[l.emission_id].radiance_node, ip.ray.wvl, gp.u, gp.v);
		o.posPDF	= pdfA;
		o.posIsArea = true;
		o.lightPos	= pos;
	}

	RayS nextRay(const IP& ip, V3 d, uint32_t rayFlags, float minT, float maxT) const
	{ // IntersectionPoint::nextRay :116-124 + Ray::next, Ray.h:102-122
		const V3 oN = dot(d, ip.N) < 0 ? -ip.N : ip.N;
		RayS r		= ip.ray;
		r.O			= safePosition(ip.P, d, oN);
		r.D			= d;
		r.depth += 1;
		r.tmin = minT;
		r.tmax = maxT;
		r.flags |= rayFlags;
		return r;
	}

	void handleNEE(const IP& ip, uint32_t matID, PathState& cur, Rng& rnd)
	{ // direct.cpp:233-352
		const prb_scene_desc& d = *sc.d;
		if (d.n_lights == 0)
			return; // no selector: LightSampler::sample returns nullptr without drawing
		float selPdf;
		const int lightID = sampleDiscrete(d.light_cdf, (int)d.n_lights + 1, rnd.getFloat(), selPdf);
		if (lightID >= (int)d.n_lights)
			return;
		const prb_light& light = d.lights[lightID];
		LightSample ls;
		sampleLight(light, ip, rnd, ls);
		const float sqrD	  = norm2(ls.lightPos - ip.P);
		const V3 L			  = ls.outgoing;
		const float cosC	  = std::abs(dot(L, ip.N));
		const float cosL	  = std::abs(ls.cosLight);
		const bool isFeasible = cosC * cosL > 1e-5f && sqrD > 1e-5f; // GEOMETRY_EPS, DISTANCE_EPS
		if (!isFeasible)
			return;
		MatCtx mc;
		mc.V		= toTangentSpace(ip.N, ip.Nx, ip.Ny, -ip.ray.D);
		mc.L		= toTangentSpace(ip.N, ip.Nx, ip.Ny, L);
		mc.wvl		= ip.ray.wvl;
		mc.u		= ip.g.u;
		mc.v		= ip.g.v;
		mc.rayFlags = ip.ray.flags;
		MatEval mout;
		materialEval(sc, matID, mc, mout);
		if (mout.flags & MSF_Delta)
			return;
		const bool rayMono		  = ip.ray.flags & PRB_RAY_MONOCHROME;
		const bool bsdfMono		  = ((mout.flags & MSF_Delta) && (mout.flags & MSF_SpectralVarying)) || rayMono;
		const Blob rayHeroFactor  = rayMono ? heroOnly() : blob(1);
		const Blob heroFactor	  = bsdfMono ? heroOnly() : blob(1);
		const Blob bsdfWvlPdfS	  = mout.pdf * heroFactor;
		if (allLE(bsdfWvlPdfS, 1e-6f)) // PDF_EPS
			return;
		const Blob connectionW = ls.radiance * mout.weight;
		const bool worthACheck = !blobIsZero(connectionW, PR_EPSILON);
		float lightPdfS		   = 0;
		if (ls.delta) { // light->hasDeltaDistribution(), direct.cpp:288-289
			lightPdfS = 1;
		} else {
			if (ls.infinite) {
				lightPdfS = ls.dirPDF_S;
			} else {
				lightPdfS = ls.posPDF;
				if (ls.posIsArea)
					lightPdfS = lightPdfS * sqrD / cosL; // IS::toSolidAngle
			}
			lightPdfS *= selPdf;
			if (!std::isnormal(lightPdfS) || lightPdfS <= 1e-6f)
				return;
		}
		const Blob lightPdfS2 = rayHeroFactor * lightPdfS; // lightPdfS * Wavelength_PDF(=1) * rayHeroFactor
		if (allLE(lightPdfS2, 1e-6f))
			return;
		const bool power = d.settings.mis_power;
		Blob mis;
		if (d.settings.do_direct && !cur.LastWasEmissive) {
			const uint32_t cameraPathLength = ip.ray.depth + 1;
			const float cameraRoulette		= rrProbability(cameraPathLength, false);
			const Blob bsdfPdfS				= bsdfWvlPdfS * cameraRoulette;
			const float denom				= bsum(misTerm(power, cur.PathPDF * lightPdfS2)) + bsum(misTerm(power, cur.PathPDF * bsdfPdfS));
			mis = ls.delta ? heroFactor / bsum(heroFactor)
						   : blob(misTerm(power, cur.PathPDF[0] * lightPdfS2[0])) / ((heroFactor * denom) * misTerm(power, cur.WavelengthPDF));
		} else {
			mis = heroFactor / (cur.WavelengthPDF * bsum(heroFactor));
		}
		const float distance = ls.infinite ? PR_INF : std::sqrt(sqrD);
		const RayS shadow	 = nextRay(ip, L, PRB_RAY_SHADOW, 0.0001f, distance);
		bool isVisible		 = false;
		if (worthACheck) {
			stats.c[S_SHADOW]++;
			Hit h; // Scene::traceShadowRay: tnear = MinT, tfar = distance - 0.001, Scene.cpp:266-280
			isVisible = !traceScene(sc, A, shadow.O, shadow.D, shadow.tmin, distance - 0.001f, true, h);
		}
		const Blob contrib = isVisible ? connectionW / lightPdfS2[0] : blob(0);
		if (ls.infinite)
			stats.c[S_BG_HIT]++;
		else
			stats.c[S_ENTITY_HIT]++;
		logState(FK_NEE, (bsdfMono ? FF_BSDF_MONO : 0) | (ls.delta ? FF_LIGHT_DELTA : 0) | (isVisible ? FF_VISIBLE : 0) | (ls.infinite ? FF_LIGHT_INFINITE : 0), ip.ray.depth, cur);
		if (log) {
			pending.lightPdfS = lightPdfS;
			pending.roulette  = rrProbability(ip.ray.depth + 1, false);
			for (int i = 0; i < 4; ++i)
				pending.bsdfPDF[i] = mout.pdf[i];
		}
		path.push_back(scatterToken(mout.type)); // direct.cpp:337-351
		path.push_back(ls.infinite ? TOK_BACKGROUND : TOK_EMISSIVE);
		pushSpectralFragment(mis, cur.Throughput, contrib, shadow.flags);
		path.pop_back();
		path.pop_back();
	}

	bool handleScattering(const IP& ip, uint32_t matID, PathState& cur, Rng& rnd, RayS& next)
	{ // direct.cpp:170-230
		const prb_scene_desc& d = *sc.d;
		cur.LastPosition		= ip.P;
		cur.LastNormal			= ip.N;
		const bool onlyDelta	= matID < d.n_materials && (d.materials[matID].flags & PRB_MATF_ONLY_DELTA);
		const float scatProb	= rrProbability(ip.ray.depth + 1, onlyDelta);
		if (scatProb <= PR_EPSILON)
			return false;
		if (scatProb < 1.0f) {
			if (rnd.getFloat() > scatProb)
				return false;
		}
		if (matID >= d.n_materials)
			return false;
		MatCtx mc;
		mc.V		= toTangentSpace(ip.N, ip.Nx, ip.Ny, -ip.ray.D);
		mc.L		= mk(0, 0, 0);
		mc.wvl		= ip.ray.wvl;
		mc.u		= ip.g.u;
		mc.v		= ip.g.v;
		mc.rayFlags = ip.ray.flags;
		MatSample sout;
		materialSample(sc, matID, mc, rnd, sout);
		path.push_back(scatterToken(sout.type)); // mCameraPath.addToken(sout.Type), direct.cpp:197
		const V3 L			= normalized(fromTangentSpace(ip.N, ip.Nx, ip.Ny, sout.L)); // MaterialSampleOutput::globalL
		cur.LastWasDelta	= sout.isDelta();
		cur.PrevPathPDF		= cur.PathPDF;
		cur.PathPDF			= cur.PathPDF * (sout.pdf * scatProb);
		if (allLE(cur.PathPDF, 1e-6f))
			return false;
		cur.Throughput = cur.Throughput * sout.weight;
		if (sout.isHeroCollapsing()) {
			cur.Throughput = cur.Throughput * heroOnly();
			cur.PathPDF	   = cur.PathPDF * heroOnly();
		}
		if (blobIsZero(cur.Throughput, PR_EPSILON))
			return false;
		uint32_t rflags = PRB_RAY_BOUNCE;
		if (sout.isHeroCollapsing())
			rflags |= PRB_RAY_MONOCHROME;
		next = nextRay(ip, L, rflags, 0.0001f, PR_INF);
		return true;
	}

	bool handleCameraVertex(const IP& ip, PathState& cur, Rng& rnd, RayS& next)
	{ // direct.cpp:73-105
		const prb_scene_desc& d = *sc.d;
		const uint32_t pathLength = ip.ray.depth + 1;
		stats.c[S_ENTITY_HIT]++;
		stats.c[S_CAMERA_DEPTH]++;
		if (pathLength == 1) { // pushSPFragment -> commitShadingPoints, LocalFrameOutputDevice.cpp:252-302
			film.sampleCount[pixelIndex] += 1;
			if (film.aov) {
				float* a = film.aov + 10 * (size_t)pixelIndex;
				a[0] += ip.N.x;
				a[1] += ip.N.y;
				a[2] += ip.N.z;
				a[3] += ip.P.x;
				a[4] += ip.P.y;
				a[5] += ip.P.z;
				a[6] += ip.g.u;
				a[7] += ip.g.v;
				a[8] += std::sqrt(ip.depth2);
				a[9] += (float)ip.g.entity;
			}
			if (film.aovExt) { // BLEND_3D(AOV_Tangent / AOV_Bitangent / AOV_View), BLEND_1D(AOV_MaterialID / AOV_EmissionID), LocalFrameOutputDevice.cpp:268-284
				float* b = film.aovExt + PRB_AOV_EXT * (size_t)pixelIndex;
				b[0] += ip.Nx.x;
				b[1] += ip.Nx.y;
				b[2] += ip.Nx.z;
				b[3] += ip.Ny.x;
				b[4] += ip.Ny.y;
				b[5] += ip.Ny.z;
				b[6] += ip.ray.D.x;
				b[7] += ip.ray.D.y;
				b[8] += ip.ray.D.z;
				b[9] += (float)ip.g.material;
				b[10] += (float)ip.g.emission;
			}
		}
		const bool hasEmission = ip.g.emission != PRB_INVALID_ID;
		if (d.settings.do_direct && hasEmission) {
			handleDirectHit(ip, cur);
			if (!d.settings.emissive_scatter)
				return false;
		}
		const uint32_t matID = ip.g.material;
		if (matID >= d.n_materials)
			return false;
		const bool onlyDelta = d.materials[matID].flags & PRB_MATF_ONLY_DELTA;
		if (d.settings.do_nee && !onlyDelta && !hasEmission)
			handleNEE(ip, matID, cur, rnd);
		cur.LastWasEmissive = hasEmission;
		return handleScattering(ip, matID, cur, rnd, next);
	}

	void handleMiss(const RayS& ray, PathState& cur)
	{ // direct.cpp:415-464
		const prb_scene_desc& d = *sc.d;
		stats.c[S_BG_HIT]++;
		const bool mono		  = ray.flags & PRB_RAY_MONOCHROME;
		const Blob heroFactor = mono ? heroOnly() : blob(1);
		bool hasInf			  = false;
		for (uint32_t i = 0; i < d.n_lights; ++i)
			if (isInfLight(d.lights[i])) // HasInfLights = infiniteLightCount() != 0 (delta lights included), direct.cpp:487
				hasInf = true;
		if (!hasInf || !d.settings.do_direct) { // handleZero
			logState(FK_ZERO, 0, ray.depth, cur);
			pushSpectralFragment(heroFactor / (cur.WavelengthPDF * bsum(heroFactor)), cur.Throughput, blob(0), ray.flags);
			return;
		}
		const bool power = d.settings.mis_power;
		float denom_mis	 = 0;
		Blob radiance	 = blob(0);
		logState(FK_INF_LIGHT, 0, ray.depth, cur);
		int nInf = 0;
		for (uint32_t i = 0; i < d.n_lights; ++i) {
			const prb_light& l = d.lights[i];
			if (!isInfLight(l) || isDeltaLight(l))
				continue;
			Blob rad;
			float pdfS;
			infLightEval(l, ray, rad, pdfS);
			const float pdf_S = pdfS * l.select_pdf;
			radiance		  = radiance + rad;
			denom_mis += bsum(misTerm(power, cur.PrevPathPDF * pdf_S));
			if (nInf < 4)
				pending.infPdfS[nInf] = pdf_S;
			++nInf;
		}
		pending.extra = (float)nInf;
		if (!d.settings.do_nee || cur.LastWasDelta) {
			pushSpectralFragment(heroFactor / (cur.WavelengthPDF * bsum(heroFactor)), cur.Throughput, radiance, ray.flags);
			return;
		}
		const float denom = bsum(misTerm(power, cur.PathPDF)) + denom_mis;
		const Blob mis	  = (heroFactor * misTerm(power, cur.PathPDF[0])) / (misTerm(power, cur.WavelengthPDF) * denom);
		pushSpectralFragment(mis, cur.Throughput, radiance, ray.flags);
	}
	void envEval(const prb_light& l, const RayS& ray, Blob& rad, float& pdfS) const
	{ // EnvironmentLight::eval, environment.cpp (no distribution)
		const V3 ld = m3mul(l.inv_normal_matrix, ray.D);
		float u, v;
		uv_from_normal(ld, u, v);
		const uint32_t node = (l.env_split && ray.depth == 0) ? l.background_node : l.radiance_node;
		rad					= evalNode(sc, node, ray.wvl, u, v);
		if (l.dist_w) { // UseDistribution, environment.cpp:68-72
			pdfS				 = dist2DContinuousPdf(l, u, v);
			const float sinTheta = cr_sin(v * PR_PI);
			const float denom	 = 2 * PR_PI * PR_PI * sinTheta;
			pdfS *= (denom <= PR_EPSILON) ? 0.0f : 1.0f / denom;
		} else {
			pdfS = cos_hemi_pdf(std::abs(ld.z));
		}
	}

	void makeIP(const RayS& ray, const Hit& h, IP& ip)
	{ // RenderTileSession::traceSingleRay :80-101 + IntersectionPoint::setForSurface :61-75
		const V3 P = ray.O + h.t * ray.D;
		provideGeometryPoint(sc, h.entity, h.prim, h.u, h.v, P, ip.g);
		ip.P	  = P;
		ip.ray	  = ray;
		ip.depth2 = norm2(ray.O - P);
		ip.NdotV  = dot(ray.D, ip.g.N);
		ip.N	  = ip.g.N;
		ip.Nx	  = ip.g.Nx;
		ip.Ny	  = ip.g.Ny;
	}

	void renderSample(uint32_t px, uint32_t py, uint32_t iteration, Rng& rnd)
	{
		const prb_scene_desc& d = *sc.d;
		stats.c[S_PIXEL_SAMPLE]++;
		path.assign(1, (uint8_t)TOK_CAMERA); // mCameraPath = [Camera] (direct.cpp:67; popTokenUntil(1) after every camera path, :137)
		CameraSampleOut cs;
		constructCameraRay(sc, px + d.settings.view_x * 0, py, iteration, rnd, cs);
		grp.importance	= cs.importance;
		grp.wvl			= cs.wvl;
		grp.wvlPDF		= cs.wvlPDF;
		grp.blendWeight = cs.blendWeight;
		RayS ray;
		ray.O	  = cs.origin;
		ray.D	  = cs.dir;
		ray.tmin  = cs.tmin;
		ray.tmax  = cs.tmax;
		ray.wvl	  = cs.wvl;
		ray.depth = 0;
		ray.flags = (cs.mono ? PRB_RAY_MONOCHROME : 0) | PRB_RAY_CAMERA;
		stats.c[S_CAMERA_RAY]++;
		stats.c[S_PRIMARY]++;
		Hit h;
		const bool hit = traceScene(sc, A, ray.O, ray.D, ray.tmin, ray.tmax, false, h);
		ray.D		   = normalized(ray.D); // RayStream::getRay re-normalises on read, RayStream.cpp:167
		if (!hit) { // IntegratorUtils::handleBackgroundGroup, IntegratorUtils.h:16-53
			stats.c[S_CAMERA_DEPTH]++;
			stats.c[S_BG_HIT]++;
			path.push_back(TOK_BACKGROUND); // LightPath::createCB(), IntegratorUtils.h:19
			bool illuminated = false;
			for (uint32_t i = 0; i < d.n_lights; ++i) {
				const prb_light& l = d.lights[i];
				if (!isInfLight(l) || isDeltaLight(l))
					continue;
				illuminated = true;
				Blob rad;
				float pdfS;
				infLightEval(l, ray, rad, pdfS);
				logState(FK_BACKGROUND, 0, 0, PathState{});
				pushSpectralFragment(blob(1), blob(1), rad, ray.flags);
			}
			if (!illuminated)
				pushSpectralFragment(blob(1), blob(1), blob(0), ray.flags);
			return;
		}
		PathState cur;
		cur.WavelengthPDF = grp.wvlPDF;
		IP ip;
		makeIP(ray, h, ip);
		RayS next;
		if (!handleCameraVertex(ip, cur, rnd, next))
			return;
		ray = next;
		// Walker::traverse, vcm/Walker.h:24-44
		for (uint32_t j = ray.depth; j < d.settings.max_ray_depth; ++j) {
			stats.c[S_CAMERA_RAY]++; // RenderTileSession::traceSingleRay :63-78
			stats.c[S_BOUNCE]++;
			if (ray.flags & PRB_RAY_MONOCHROME)
				stats.c[S_MONO]++;
			Hit bh;
			if (!traceScene(sc, A, ray.O, ray.D, ray.tmin, ray.tmax, false, bh)) {
				path.push_back(TOK_BACKGROUND); // direct.cpp:124-134
				handleMiss(ray, cur);
				path.pop_back();
				break;
			}
			makeIP(ray, bh, ip);
			if (!handleCameraVertex(ip, cur, rnd, next))
				break;
			ray = next;
		}
	}
};

struct FtzScope { // setupFloatingPointEnvironment (FTZ|DAZ), src/base/Platform.h:20-34; restored on scope exit
	unsigned int saved;
	FtzScope()
		: saved(_mm_getcsr())
	{
		_MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
		_MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
	}
	~FtzScope() { _mm_setcsr(saved); }
};
#define setFTZ() FtzScope ftz_scope_

void initScene(Scene& sc, const prb_scene_desc* d)
{
	sc.d = d;
	// RussianRoulette(minDepth = soft_max_ray_depth, factor 0.9): min(1, pow(0.9f, len - min)) as float, <=1e-4 -> 0
	const uint32_t soft = d->settings.soft_max_ray_depth;
	const uint32_t n	= std::max(d->settings.max_ray_depth, soft) + 4;
	sc.rrProb.assign(n + 1, 1.0f);
	for (uint32_t len = 0; len <= n; ++len) {
		if (len >= soft) {
			const float p  = std::min<float>(1.0f, (float)std::pow((double)0.9f, (double)(len - soft)));
			sc.rrProb[len] = p <= 1e-4f ? 0.0f : p;
		}
	}
}
} // namespace

// ====================================================================================== C interface
extern "C" {
struct orc_scene {
	Scene sc;
	Accel accel;
};
orc_scene* orc_scene_create(const prb_scene_desc* d)
{
	setFTZ();
	auto* s = new orc_scene();
	initScene(s->sc, d);
	buildAccel(s->sc, s->accel);
	return s;
}
void orc_scene_destroy(orc_scene* s) { delete s; }

// Render iterations [first, first+count) of the given tiles.  rng: W*H states (updated in place).
// film_mean: W*H*3 running mean (unfiltered; updated), sample_count: W*H, aov: W*H*10 or NULL, stats: 11 counters,
// feedback: W*H words (OR of PRB_FEEDBACK_* bits, updated) or NULL; online_mean / online_variance: W*H*3 each or NULL.
// lpe_mean: n_lpe films of W*H*3 floats (running means of the light path expression channels, updated) or NULL;
// aov_ext: W*H*PRB_AOV_EXT floats (sums, updated) or NULL.
void orc_render_lpe(orc_scene* s, uint64_t* rng, const prb_tile* tiles, size_t n_tiles, uint32_t first_iteration, uint32_t iteration_count,
					float* film_mean, uint32_t* sample_count, float* aov, uint64_t* stats11, int threads, uint32_t* feedback, float* online_mean,
					float* online_variance, float* lpe_mean, float* aov_ext);
void orc_render(orc_scene* s, uint64_t* rng, const prb_tile* tiles, size_t n_tiles, uint32_t first_iteration, uint32_t iteration_count,
				float* film_mean, uint32_t* sample_count, float* aov, uint64_t* stats11, int threads, uint32_t* feedback, float* online_mean,
				float* online_variance)
{
	orc_render_lpe(s, rng, tiles, n_tiles, first_iteration, iteration_count, film_mean, sample_count, aov, stats11, threads, feedback, online_mean, online_variance,
				   nullptr, nullptr);
}
void orc_render_lpe(orc_scene* s, uint64_t* rng, const prb_tile* tiles, size_t n_tiles, uint32_t first_iteration, uint32_t iteration_count,
					float* film_mean, uint32_t* sample_count, float* aov, uint64_t* stats11, int threads, uint32_t* feedback, float* online_mean,
					float* online_variance, float* lpe_mean, float* aov_ext)
{
	const prb_settings& st = s->sc.d->settings;
	const uint32_t W	   = st.film_width;
	Film film;
	film.iterXYZ.assign((size_t)W * st.film_height * 3, 0.0f);
	film.mean		 = film_mean;
	film.sampleCount = sample_count;
	film.aov		 = aov;
	film.aovExt		 = aov_ext;
	film.feedback	 = feedback;
	film.varMean	 = online_mean;
	film.varVar		 = online_variance;
	const size_t lpeStride = (size_t)W * st.film_height * 3;
	const uint32_t nLPE	   = lpe_mean ? s->sc.d->n_lpe : 0;
	if (nLPE) {
		film.lpeMean = lpe_mean;
		film.lpeIter.assign(lpeStride * nLPE, 0.0f);
	}
	// pixel list (pixels are independent: own RNG stream, own film cell)
	std::vector<uint32_t> pixels;
	for (size_t t = 0; t < n_tiles; ++t)
		for (uint32_t y = tiles[t].sy; y < tiles[t].ey; ++y)
			for (uint32_t x = tiles[t].sx; x < tiles[t].ex; ++x)
				pixels.push_back(y * W + x);
	threads = std::max(1, threads);
	std::vector<Stats> tstats(threads);
	std::atomic<size_t> cursor{ 0 };
	auto worker = [&](int tid) {
		setFTZ();
		Integrator I{ s->sc, s->accel, film, tstats[tid], 0, Group{} };
		for (;;) {
			const size_t begin = cursor.fetch_add(256);
			if (begin >= pixels.size())
				break;
			const size_t end = std::min(pixels.size(), begin + 256);
			for (size_t i = begin; i < end; ++i) {
				const uint32_t p = pixels[i];
				Rng rnd{ rng[p] };
				for (uint32_t it = first_iteration; it < first_iteration + iteration_count; ++it) {
					I.pixelIndex = p;
					for (int c = 0; c < 3; ++c)
						film.iterXYZ[3 * (size_t)p + c] = 0.0f;
					for (uint32_t k = 0; k < nLPE; ++k)
						for (int c = 0; c < 3; ++c)
							film.lpeIter[k * lpeStride + 3 * (size_t)p + c] = 0.0f;
					I.renderSample(p % W, p / W, it, rnd);
					// FrameOutputDevice::onEndOfIteration: (a * (iteration - 1) + b) / iteration, 1-based
					const float iter = (float)(it + 1);
					for (int c = 0; c < 3; ++c) {
						float& a = film.mean[3 * (size_t)p + c];
						a		 = (a * (float)it + film.iterXYZ[3 * (size_t)p + c]) / iter;
					}
					for (uint32_t k = 0; k < nLPE; ++k) // the expression channels merge like the main one, FrameOutputDevice.cpp:216-218
						for (int c = 0; c < 3; ++c) {
							float& a = film.lpeMean[k * lpeStride + 3 * (size_t)p + c];
							a		 = (a * (float)it + film.lpeIter[k * lpeStride + 3 * (size_t)p + c]) / iter;
						}
					if (film.varMean && film.varVar) // VarianceEstimator::addValue, src/core/buffer/VarianceEstimator.inl:16-28
						for (int c = 0; c < 3; ++c) {
							const float value = film.iterXYZ[3 * (size_t)p + c];
							float& mean		  = film.varMean[3 * (size_t)p + c];
							float& var		  = film.varVar[3 * (size_t)p + c];
							const float delta = value - mean;
							mean += delta / iter;
							const float delta2 = value - mean;
							var				   = (var * (float)it + delta * delta2) / iter;
						}
				}
				rng[p] = rnd.s;
			}
		}
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < threads; ++t)
		pool.emplace_back(worker, t);
	worker(0);
	for (auto& th : pool)
		th.join();
	if (stats11)
		for (int t = 0; t < threads; ++t)
			for (int i = 0; i < 11; ++i)
				stats11[i] += tstats[t].c[i];
}

// Test instrumentation: renders iterations [first, first + count) of the listed pixels single-threaded WITHOUT touching a
// film and returns the fragment log (FragLog records of 48 floats); returns the number of records (<= capacity kept).
size_t orc_log_fragments(orc_scene* s, const uint64_t* rng, const uint32_t* pixels, size_t n_pixels, uint32_t first_iteration, uint32_t iteration_count,
						 float* out48, size_t capacity)
{
	setFTZ();
	const prb_settings& st = s->sc.d->settings;
	const uint32_t W	   = st.film_width;
	Film film;
	film.iterXYZ.assign((size_t)W * st.film_height * 3, 0.0f);
	std::vector<float> mean((size_t)W * st.film_height * 3, 0.0f);
	std::vector<uint32_t> cnt((size_t)W * st.film_height, 0);
	film.mean		 = mean.data();
	film.sampleCount = cnt.data();
	film.aov		 = nullptr;
	Stats stats;
	std::vector<FragLog> log;
	Integrator I{ s->sc, s->accel, film, stats, 0, Group{} };
	I.log = &log;
	for (size_t i = 0; i < n_pixels; ++i) {
		const uint32_t p = pixels[i];
		Rng rnd{ rng[p] };
		for (uint32_t it = first_iteration; it < first_iteration + iteration_count; ++it) {
			I.pixelIndex = p;
			I.renderSample(p % W, p / W, it, rnd);
		}
	}
	const size_t n = std::min(log.size(), capacity);
	std::memcpy(out48, log.data(), n * sizeof(FragLog));
	return log.size();
}

// pixel filter as a post pass (zero padded convolution with the FilterCache table)
void orc_apply_filter(const prb_scene_desc* d, const float* in_xyz, float* out_xyz)
{
	const int r = d->settings.filter_radius, W = (int)d->settings.film_width, H = (int)d->settings.film_height;
	const float* tab = d->pool + d->settings.filter_offset;
	const int dia	 = 2 * r + 1;
	for (size_t i = 0; i < (size_t)W * H * 3; ++i)
		out_xyz[i] = 0;
	for (int y = 0; y < H; ++y)
		for (int x = 0; x < W; ++x)
			for (int dy = -r; dy <= r; ++dy)
				for (int dx = -r; dx <= r; ++dx) {
					const int sx = x + dx, sy = y + dy;
					if (sx < 0 || sy < 0 || sx >= W || sy >= H)
						continue;
					const float w = tab[(dy + r) * dia + (dx + r)];
					if (!(w > PR_EPSILON))
						continue;
					for (int c = 0; c < 3; ++c)
						out_xyz[3 * ((size_t)sy * W + sx) + c] += w * in_xyz[3 * ((size_t)y * W + x) + c];
				}
}

void orc_trace_closest(orc_scene* s, const prb_ray_soa* rays, size_t n, prb_hit_soa* hits, int threads)
{
	threads = std::max(1, threads);
	std::vector<std::thread> pool;
	auto work = [&](int tid) {
		setFTZ();
		for (size_t i = tid; i < n; i += threads) {
			Hit h;
			const V3 O = mk(rays->org_x[i], rays->org_y[i], rays->org_z[i]), D = mk(rays->dir_x[i], rays->dir_y[i], rays->dir_z[i]);
			const float tmin = rays->tmin ? rays->tmin[i] : 0.0001f, tmax = rays->tmax ? rays->tmax[i] : PR_INF;
			const bool ok		  = traceScene(s->sc, s->accel, O, D, tmin, tmax, false, h);
			hits->entity_id[i]	  = ok ? h.entity : PRB_INVALID_ID;
			hits->primitive_id[i] = ok ? h.prim : PRB_INVALID_ID;
			hits->u[i]			  = ok ? h.u : 0;
			hits->v[i]			  = ok ? h.v : 0;
			hits->t[i]			  = ok ? h.t : tmax;
		}
	};
	for (int t = 1; t < threads; ++t)
		pool.emplace_back(work, t);
	work(0);
	for (auto& th : pool)
		th.join();
}
void orc_trace_any(orc_scene* s, const prb_ray_soa* rays, size_t n, uint8_t* occluded, int threads)
{
	threads = std::max(1, threads);
	std::vector<std::thread> pool;
	auto work = [&](int tid) {
		setFTZ();
		for (size_t i = tid; i < n; i += threads) {
			Hit h;
			const V3 O = mk(rays->org_x[i], rays->org_y[i], rays->org_z[i]), D = mk(rays->dir_x[i], rays->dir_y[i], rays->dir_z[i]);
			const float tmin = rays->tmin ? rays->tmin[i] : 0.0001f, tmax = rays->tmax ? rays->tmax[i] : PR_INF;
			occluded[i] = traceScene(s->sc, s->accel, O, D, tmin, tmax, true, h) ? 1 : 0;
		}
	};
	for (int t = 1; t < threads; ++t)
		pool.emplace_back(work, t);
	work(0);
	for (auto& th : pool)
		th.join();
}

// camera rays of one iteration (does not modify rng)
size_t orc_generate_camera_rays(orc_scene* s, const uint64_t* rng, const prb_tile* tiles, size_t n_tiles, uint32_t iteration, float* org_xyz,
								float* dir_xyz, float* wavelengths4, uint32_t* pixel_index, size_t capacity)
{
	setFTZ();
	const uint32_t W = s->sc.d->settings.film_width;
	size_t n		 = 0;
	for (size_t t = 0; t < n_tiles; ++t)
		for (uint32_t y = tiles[t].sy; y < tiles[t].ey; ++y)
			for (uint32_t x = tiles[t].sx; x < tiles[t].ex; ++x) {
				if (n >= capacity)
					return n;
				Rng rnd{ rng[y * W + x] };
				CameraSampleOut cs;
				constructCameraRay(s->sc, x, y, iteration, rnd, cs);
				org_xyz[3 * n]	   = cs.origin.x;
				org_xyz[3 * n + 1] = cs.origin.y;
				org_xyz[3 * n + 2] = cs.origin.z;
				dir_xyz[3 * n]	   = cs.dir.x;
				dir_xyz[3 * n + 1] = cs.dir.y;
				dir_xyz[3 * n + 2] = cs.dir.z;
				for (int i = 0; i < 4; ++i)
					wavelengths4[4 * n + i] = cs.wvl[i];
				pixel_index[n] = y * W + x;
				++n;
			}
	return n;
}

static MatCtx ctxOf(const prb_material_query& q)
{
	MatCtx c;
	c.V = mk(q.V[0], q.V[1], q.V[2]);
	c.L = mk(q.L[0], q.L[1], q.L[2]);
	for (int i = 0; i < 4; ++i)
		c.wvl[i] = q.wavelength_nm[i];
	c.u		   = q.uv[0];
	c.v		   = q.uv[1];
	c.rayFlags = q.ray_flags;
	return c;
}
// Light::sample (NEE form) of light `light_id` for a point P, followed by IInfiniteLight::eval in the sampled direction.
// out: outgoing[3], sampled Direction_PDF_S, sampled radiance[4], evaluated Direction_PDF_S, evaluated radiance[4], delta flag.
void orc_light_sample_and_eval(orc_scene* s, uint32_t light_id, const float* P, const float* wvl4, uint64_t* rng_state, float* out14)
{
	setFTZ();
	Film film{};
	Stats stats;
	Integrator in{ s->sc, s->accel, film, stats, 0, Group{} };
	IP ip{};
	ip.P = ld3(P);
	for (int k = 0; k < 4; ++k)
		ip.ray.wvl[k] = wvl4[k];
	Rng rnd{ *rng_state };
	Integrator::LightSample ls;
	const prb_light& l = s->sc.d->lights[light_id];
	in.sampleLight(l, ip, rnd, ls);
	*rng_state = rnd.s;
	out14[0] = ls.outgoing.x, out14[1] = ls.outgoing.y, out14[2] = ls.outgoing.z;
	out14[3] = ls.dirPDF_S;
	for (int k = 0; k < 4; ++k)
		out14[4 + k] = ls.radiance[k];
	RayS ray{};
	ray.D	= ls.outgoing;
	ray.wvl = ip.ray.wvl;
	Blob rad = blob(0);
	float pdf = 0;
	if (l.type != PRB_LIGHT_AREA && l.type != PRB_LIGHT_SUN_DELTA)
		in.infLightEval(l, ray, rad, pdf);
	out14[8] = pdf;
	for (int k = 0; k < 4; ++k)
		out14[9 + k] = rad[k];
	out14[13] = ls.delta ? 1.0f : 0.0f;
}
void orc_material_eval(orc_scene* s, const prb_material_query* q, size_t n, prb_material_result* out)
{
	setFTZ();
	for (size_t i = 0; i < n; ++i) {
		MatEval e;
		materialEval(s->sc, q[i].material_id, ctxOf(q[i]), e);
		for (int k = 0; k < 4; ++k) {
			out[i].weight[k] = e.weight[k];
			out[i].pdf_s[k]	 = e.pdf[k];
		}
		out[i].L[0] = out[i].L[1] = out[i].L[2] = 0;
		out[i].flags							= e.flags;
		out[i].type								= e.type;
		out[i].rng_state						= q[i].rng_state;
	}
}
void orc_material_sample(orc_scene* s, const prb_material_query* q, size_t n, prb_material_result* out)
{
	setFTZ();
	for (size_t i = 0; i < n; ++i) {
		MatSample e;
		Rng rnd{ q[i].rng_state };
		materialSample(s->sc, q[i].material_id, ctxOf(q[i]), rnd, e);
		for (int k = 0; k < 4; ++k) {
			out[i].weight[k] = e.weight[k];
			out[i].pdf_s[k]	 = e.pdf[k];
		}
		out[i].L[0]		 = e.L.x;
		out[i].L[1]		 = e.L.y;
		out[i].L[2]		 = e.L.z;
		out[i].flags	 = e.flags;
		out[i].type		 = e.type;
		out[i].rng_state = rnd.s;
	}
}

// ---- unit-level probes for the known-answer tests ported from the reference's src/tests
float orc_fresnel_dielectric(float cosI, float n_in, float n_out) { return fresnel_dielectric(cosI, n_in, n_out); }
float orc_fresnel_conductor(float cosI, float n_in, float n_out, float k) { return fresnel_conductor(cosI, n_in, n_out, k); }
float orc_fresnel_schlick(float d, float n1, float n2) { return schlick3(d, n1, n2); }
float orc_ndf_ggx_iso(const float* H, float r) { return ndf_ggx(mk(H[0], H[1], H[2]), r); }
float orc_ndf_ggx_aniso(const float* H, float rx, float ry) { return ndf_ggx(mk(H[0], H[1], H[2]), rx, ry); }
float orc_pdf_ggx_iso(const float* H, float r) { return pdf_ggx(mk(H[0], H[1], H[2]), r); }
float orc_pdf_ggx_aniso(const float* H, float rx, float ry) { return pdf_ggx(mk(H[0], H[1], H[2]), rx, ry); }
float orc_microfacet_reflection_eval_conductor(const float* wIn, const float* wOut, float m1, float m2, int aniso, int vndf, float ior, float kappa)
{
	MicrofacetReflection r{ RoughDistribution{ m1, m2, aniso != 0, vndf != 0 } };
	return r.evalConductor(mk(wIn[0], wIn[1], wIn[2]), mk(wOut[0], wOut[1], wOut[2]), ior, kappa);
}
void orc_reflect(const float* V, const float* N, float* out)
{
	const V3 r = reflect(mk(V[0], V[1], V[2]), mk(N[0], N[1], N[2]));
	out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
void orc_refract(float eta, const float* V, const float* N, float* out, int* total)
{
	bool t;
	const V3 r = refract(eta, mk(V[0], V[1], V[2]), mk(N[0], N[1], N[2]), t);
	out[0] = r.x, out[1] = r.y, out[2] = r.z;
	*total = t;
}
void orc_halfway_reflection(const float* a, const float* b, float* out)
{
	const V3 r = halfway_reflection(mk(a[0], a[1], a[2]), mk(b[0], b[1], b[2]));
	out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
void orc_cos_hemi(float u1, float u2, float* out)
{
	const V3 r = cos_hemi(u1, u2);
	out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
void orc_tangent_frame(const float* N, float* Nx, float* Ny)
{
	V3 x, y;
	tangent_frame(mk(N[0], N[1], N[2]), x, y);
	Nx[0] = x.x, Nx[1] = x.y, Nx[2] = x.z;
	Ny[0] = y.x, Ny[1] = y.y, Ny[2] = y.z;
}
// Spherical::cartesian_from_uv / uv_from_normal (src/base/math/Spherical.h), for the known answers of src/tests/sphere.cpp
void orc_cartesian_from_uv(float u, float v, float* out)
{
	const V3 n = cartesian_from_uv(u, v);
	out[0] = n.x, out[1] = n.y, out[2] = n.z;
}
void orc_uv_from_normal(const float* N, float* uv) { uv_from_normal(mk(N[0], N[1], N[2]), uv[0], uv[1]); }
void orc_random_stream(uint64_t seed, uint32_t n, uint32_t* out32, float* outf)
{
	Rng a{ seed | 3u }, b{ seed | 3u };
	for (uint32_t i = 0; i < n; ++i) {
		if (out32)
			out32[i] = a.get32();
		if (outf)
			outf[i] = b.getFloat();
	}
}
float orc_eval_node(orc_scene* s, uint32_t node, float wavelength, float u, float v) { return evalNode(s->sc, node, blob(wavelength), u, v)[0]; }
// MicrofacetReflection<...>::eval / ::pdf (src/base/math/MicrofacetReflection.h), src/tests/microfacets.cpp reciprocity cases
float orc_microfacet_reflection_eval(const float* wIn, const float* wOut, float m1, float m2, int aniso, int vndf)
{
	MicrofacetReflection r{ RoughDistribution{ m1, m2, aniso != 0, vndf != 0 } };
	return r.eval(mk(wIn[0], wIn[1], wIn[2]), mk(wOut[0], wOut[1], wOut[2]));
}
float orc_microfacet_reflection_pdf(const float* wIn, const float* wOut, float m1, float m2, int aniso, int vndf)
{
	MicrofacetReflection r{ RoughDistribution{ m1, m2, aniso != 0, vndf != 0 } };
	return r.pdf(mk(wIn[0], wIn[1], wIn[2]), mk(wOut[0], wOut[1], wOut[2]));
}
void orc_halfway_refractive(float n_in, const float* a, float n_out, const float* b, float* out)
{
	const V3 r = halfway_refractive(n_in, mk(a[0], a[1], a[2]), n_out, mk(b[0], b[1], b[2]));
	out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
// Distribution1D over a caller-built CDF (size entries, cdf[0] = 0, cdf[size-1] = 1): src/tests/distribution.cpp
float orc_sample_continuous(const float* cdf, int size, float u, float* pdf) { return sampleContinuous(cdf, size, u, *pdf); }
int orc_sample_discrete(const float* cdf, int size, float u, float* pdf) { return sampleDiscrete(cdf, size, u, *pdf); }
}
