// Device shading: node evaluation, geometry points, the six materials, samplers / spectral mappers, camera
// rays, light sampling.  Each function cites the reference file:line it implements.  Everything works on the
// 4-wavelength SpectralBlob held in registers; no tensor cores (the path is not a dense contraction).
#pragma once
#include "dev_bvh.cuh"

namespace prb {
// ------------------------------------------------------------------ spectra / nodes
PRB_DEV_MED float tableLookup(const float* data, uint32_t count, float start, float end, float w)
{ // EquidistantSpectrumView::lookup, src/core/spectral/EquidistantSpectrum.inl:34-41
	const float delta = (end - start) / (count - 1);
	const float af	  = fmaxf(0.0f, (w - start) / delta);
	const int index	  = (int)fminf((float)(count - 2), af);
	const float t	  = fminf((float)(count - 1), af) - index;
	return __ldg(data + index) * (1 - t) + __ldg(data + index + 1) * t;
}
constexpr float CIE_START = 390, CIE_END = 830, CIE_RANGE = CIE_END - CIE_START;
constexpr int CIE_N		   = 441;
constexpr float CIE_Y_NORM = 113.042314572337f * (CIE_RANGE / (CIE_N - 1));
PRB_DEV float cieEval(const DScene& S, int c, float w)
{ // CIE::eval_x/y/z, src/core/spectral/CIE.h:41-58
	return divPositive(tableLookup(S.pool + S.cieOffset + c * CIE_N, CIE_N, CIE_START, CIE_END, w), CIE_Y_NORM) * CIE_RANGE; // z-bar is zero above 650 nm: see divPositive
}

// leaf node kinds; MUL / CHECKER reference other nodes.  The flattened graph is shallow (depth <= 3 in the
// config scenes); an explicit small stack avoids device recursion.
PRB_DEV Blob evalLeafNode(const DScene& S, const prb_node& n, const Blob& w)
{
	Blob r;
	switch (n.type) {
	default:
	case PRB_NODE_CONST: return blob(n.p[0]);
	case PRB_NODE_PARAM:
	case PRB_NODE_PARAM_SCALED: // SpectralUpsampler::compute, SpectralUpsampler.h:45-49
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const float x = (n.p[0] * w[i] + n.p[1]) * w[i] + n.p[2];
			r[i]		  = 0.5f * x * (1.0f / sqrtf(x * x + 1.0f)) + 0.5f;
			if (n.type == PRB_NODE_PARAM_SCALED)
				r[i] = r[i] * n.p[3];
		}
		return r;
	case PRB_NODE_TABLE:
#pragma unroll
		for (int i = 0; i < 4; ++i)
			r[i] = tableLookup(S.pool + n.a, n.b, n.p[0], n.p[1], w[i]);
		return r;
	case PRB_NODE_SELLMEIER: { // Scattering::sellmeier2 + sqrt, Scattering.h:219-242
		const float* B = S.pool + n.a;
		const float* C = B + n.b;
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const float qm	= fdiv(w[i], 1000.0f);
			const float qm2 = qm * qm;
			float value		= 1;
			for (uint32_t k = 0; k < n.b; ++k)
				value += __ldg(B + k) * qm2 / (qm2 - __ldg(C + k));
			r[i] = sqrtf(value);
		}
		return r;
	}
	}
}
// ---- image textures: NonParametricImageNode::eval, loader/shader/ImageNode.cpp:131-162
// texel (x, y) of an RGB image under the wrap modes of OpenImageIO's TextureOpt (black: zero outside, clamp, periodic, mirror)
PRB_DEV bool wrapTexel(int& i, int size, int mode)
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
PRB_DEV void fetchTexel(const float* img, int w, int h, int x, int y, int wrapS, int wrapT, float rgb[3])
{
	rgb[0] = rgb[1] = rgb[2] = 0;
	if (!wrapTexel(x, w, wrapS) || !wrapTexel(y, h, wrapT))
		return;
	const float* t = img + 3 * ((size_t)y * w + x);
	rgb[0] = __ldg(t), rgb[1] = __ldg(t + 1), rgb[2] = __ldg(t + 2);
}
// B-spline weights of OpenImageIO's bicubic lookup
PRB_DEV void bsplineWeights(float f, float w[4])
{
	const float one_f = 1.0f - f;
	w[0]			  = fdiv(one_f * one_f * one_f, 6.0f);
	w[1]			  = 2.0f / 3.0f - 0.5f * f * f * (2.0f - f);
	w[2]			  = 2.0f / 3.0f - 0.5f * one_f * one_f * (2.0f - one_f);
	w[3]			  = fdiv(f * f * f, 6.0f);
}
// SpectralUpsampler::prepare for one RGB triple, src/core/spectral/SpectralUpsampler.cpp (trilinear lookup in the coefficient cube)
PRB_DEV void upsamplerPrepare(const DScene& S, const float rgb[3], float coeffs[3])
{
	constexpr float EPS = 0.0001f;
	if (rgb[0] <= EPS && rgb[1] <= EPS && rgb[2] <= EPS) {
		coeffs[0] = 0, coeffs[1] = 0, coeffs[2] = -500.0f;
		return;
	}
	if (1 - rgb[0] <= EPS && 1 - rgb[1] <= EPS && 1 - rgb[2] <= EPS) {
		coeffs[0] = 0, coeffs[1] = 0, coeffs[2] = 5000000.0f;
		return;
	}
	const uint32_t res = S.upsamplerRes;
	const float* scale = S.pool + S.upsamplerOffset;
	const float* d	   = scale + res;
	const uint32_t dx = 3, dy = 3 * res, dz = 3 * res * res;
	int largest = 0;
	for (int j = 1; j < 3; ++j)
		if (rgb[largest] <= rgb[j])
			largest = j;
	const float z	 = rgb[largest];
	const float sc	 = fdiv((float)(res - 1), z);
	const float x	 = rgb[(largest + 1) % 3] * sc;
	const float y	 = rgb[(largest + 2) % 3] * sc;
	const uint32_t xi = min((uint32_t)x, res - 2);
	const uint32_t yi = min((uint32_t)y, res - 2);
	int left = 0, size = (int)res - 2; // find_interval
	const int lastInterval = (int)res - 2;
	while (size > 0) {
		const int half = size >> 1, middle = left + half + 1;
		if (__ldg(scale + middle) < z) {
			left = middle;
			size -= half + 1;
		} else {
			size = half;
		}
	}
	const uint32_t zi = (uint32_t)min(left, lastInterval);
	uint32_t off	  = (((largest * res + zi) * res + yi) * res + xi) * 3;
	const float x1 = x - (float)xi, x0 = 1.0f - x1, y1 = y - (float)yi, y0 = 1.0f - y1;
	const float z1 = fdiv(z - __ldg(scale + zi), __ldg(scale + zi + 1) - __ldg(scale + zi)), z0 = 1.0f - z1;
	for (int j = 0; j < 3; ++j) {
		coeffs[j] = ((__ldg(d + off) * x0 + __ldg(d + off + dx) * x1) * y0 + (__ldg(d + off + dy) * x0 + __ldg(d + off + dy + dx) * x1) * y1) * z0
					+ ((__ldg(d + off + dz) * x0 + __ldg(d + off + dz + dx) * x1) * y0 + (__ldg(d + off + dz + dy) * x0 + __ldg(d + off + dz + dy + dx) * x1) * y1) * z1;
		++off;
	}
}
PRB_DEV float srgbLinearize(float x)
{ // RGBConverter::linearize, src/core/spectral/RGBConverter.cpp:53-58 -- as written there, `x / 12.92 * x` on the linear segment
	if (x <= 0.04045f)
		return fdiv(x, 12.92f) * x;
	return (float)pow((double)fdiv(x + 0.055f, 1.055f), (double)2.4f);
}
PRB_DEV prb_node loadNode(const DScene& S, uint32_t id)
{ // 32 bytes as two 128-bit loads
	const uint4* p = reinterpret_cast<const uint4*>(S.nodes + id);
	const uint4 a = __ldg(p), b = __ldg(p + 1);
	prb_node n;
	n.type = a.x, n.flags = a.y, n.a = a.z, n.b = a.w;
	n.p[0] = __uint_as_float(b.x), n.p[1] = __uint_as_float(b.y), n.p[2] = __uint_as_float(b.z), n.p[3] = __uint_as_float(b.w);
	return n;
}
// (takes the node id, not the node: a by-reference argument would force the caller's copy of the node into local memory)
__device__ __noinline__ Blob evalImageNode(const DScene& S, uint32_t nodeID, const Blob& wvl, float u, float v)
{
	const prb_node n = loadNode(S, nodeID);
	const int w = (int)(n.b & 0xFFFFu), h = (int)(n.b >> 16);
	const float* img = S.pool + n.a;
	const int interp = (int)n.p[0], wrapS = (int)n.p[1], wrapT = (int)n.p[2];
	// TextureSystem::texture(file, opts, s = u, t = 1 - v, no derivatives), ImageNode.cpp:141-145; texel centres at (i + 0.5) / size
	const float x = u * (float)w - 0.5f, y = (1 - v) * (float)h - 0.5f;
	const float flx = floorf(x), fly = floorf(y);
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
	upsamplerPrepare(S, rgb, k);
	Blob r;
#pragma unroll
	for (int i = 0; i < 4; ++i) { // SpectralUpsampler::compute, SpectralUpsampler.h:45-49
		const float q = (k[0] * wvl[i] + k[1]) * wvl[i] + k[2];
		r[i]		  = 0.5f * q * (1.0f / sqrtf(q * q + 1.0f)) + 0.5f;
	}
	return r;
}
PRB_DEV bool checkerSelectsB(const prb_node& n, float u, float v)
{ // CheckerboardNode::check, CheckerboardNode.cpp:26-40
	float cu = u, cv = v;
	if (n.p[2] == 1.0f) {
		cu = u * n.p[0];
		cv = v * n.p[0];
	} else if (n.p[2] == 2.0f) {
		cu = u * n.p[0];
		cv = v * n.p[1];
	}
	return ((int)floorf(cu) + (int)floorf(cv)) % 2 == 0;
}
// Two out-of-line forms: evalNodeBase knows no image nodes and makes no calls of its own (a leaf function: no register saves
// around a call -- ncu on the Cornell box: the call to evalImageNode inside the loop doubled the local-memory stores of
// k_shade and cost 5 % although it is never taken there), evalNodeTex also fetches image textures.  The kernels that inline
// the Lambert code use evalNodeBase (MatCtx::noImageNodes; the host picks them only for scenes without image nodes).
template <bool TEX>
PRB_DEV Blob evalNodeBody(const DScene& S, uint32_t id, const Blob& w, float u, float v)
{
	// product of factors: MUL nodes push both operands, CHECKER picks one; leaves multiply into the result
	uint32_t st[8];
	int sp	 = 0;
	st[sp++] = id;
	Blob acc = blob(1.0f);
	bool first = true;
	while (sp > 0) {
		const uint32_t nodeID = st[--sp];
		const prb_node n	  = loadNode(S, nodeID);
		if (n.type == PRB_NODE_MUL) {
			if (sp + 2 <= 8) {
				st[sp++] = n.b; // evaluated second: a * b with a first
				st[sp++] = n.a;
			}
		} else if (n.type == PRB_NODE_CHECKER) {
			if (sp < 8)
				st[sp++] = checkerSelectsB(n, u, v) ? n.b : n.a;
		} else {
			Blob leaf;
			if (TEX && n.type == PRB_NODE_IMAGE)
				leaf = evalImageNode(S, nodeID, w, u, v);
			else
				leaf = evalLeafNode(S, n, w);
			acc	  = first ? leaf : acc * leaf;
			first = false;
		}
	}
	return acc;
}
__device__ __noinline__ Blob evalNodeBase(const DScene& S, uint32_t id, const Blob& w, float u, float v) { return evalNodeBody<false>(S, id, w, u, v); }
__device__ __noinline__ Blob evalNodeTex(const DScene& S, uint32_t id, const Blob& w, float u, float v) { return evalNodeBody<true>(S, id, w, u, v); }
PRB_DEV Blob evalNode(const DScene& S, uint32_t id, const Blob& w, float u, float v) { return evalNodeTex(S, id, w, u, v); }
// The kernels that inline the Lambert code (scenes without image nodes) evaluate a LEAF node -- nearly every albedo is one -- and
// the PRODUCT OF TWO LEAVES (an emitter's spectrum x its scale) in line: the node loads and the leaf formula instead of the
// out-of-line graph walk with its stack in local memory and its generic loads of the scene descriptor (ncu on the Cornell box:
// the walk for the light's radiance was 12 % of k_shade's instructions).  Same factors in the same order: same value.
PRB_DEV bool isLeafNode(const prb_node& n) { return n.type != PRB_NODE_MUL && n.type != PRB_NODE_CHECKER; }
PRB_DEV Blob evalNodeFast(const DScene& S, uint32_t id, const Blob& w, float u, float v)
{
	prb_node n	   = loadNode(S, id);
	uint32_t other = PRB_INVALID_ID; // second factor of a product of two leaves
	if (n.type == PRB_NODE_MUL) {
		const uint32_t ida = n.a, idb = n.b;
		const prb_node a = loadNode(S, ida);
		if (!isLeafNode(a) || !isLeafNode(loadNode(S, idb)))
			return evalNodeBase(S, id, w, u, v);
		n	  = a;
		other = idb;
	} else if (!isLeafNode(n)) {
		return evalNodeBase(S, id, w, u, v);
	}
	Blob acc = blob(1.0f);
#pragma unroll 1
	for (int k = 0; k < 2; ++k) { // one copy of the leaf code
		const Blob leaf = evalLeafNode(S, n, w);
		acc				= k == 0 ? leaf : acc * leaf;
		if (other == PRB_INVALID_ID)
			break;
		n	  = loadNode(S, other);
		other = PRB_INVALID_ID;
		if (k == 1)
			break;
	}
	return acc;
}

// ------------------------------------------------------------------ geometry point (GeometryPoint.h:10-25)
struct GeomPoint {
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
PRB_DEV void getFace(const DScene& S, const prb_mesh& m, uint32_t f, FaceData& fd, bool wantN, bool wantUV)
{ // MeshBase::getFace, src/core/mesh/MeshBase.inl:96-134
	const uint4 idx	 = __ldg(reinterpret_cast<const uint4*>(S.faceIndices) + (m.face_offset + f));
	const uint32_t i[4] = { idx.x, idx.y, idx.z, idx.w };
	fd.quad				= idx.w != PRB_INVALID_ID;
	const int n			= fd.quad ? 4 : 3;
	for (int j = 0; j < n; ++j) {
		fd.V[j] = ld3(S.vertices + 3 * (size_t)(m.vertex_offset + i[j]));
		if (wantN && (m.features & PRB_MESH_HAS_NORMALS))
			fd.N[j] = ld3(S.normals + 3 * (size_t)(m.normal_offset + i[j]));
		if (wantUV && (m.features & PRB_MESH_HAS_UVS)) {
			fd.UV[j][0] = S.uvs[2 * (size_t)(m.uv_offset + i[j])];
			fd.UV[j][1] = S.uvs[2 * (size_t)(m.uv_offset + i[j]) + 1];
		}
	}
	fd.slot = S.faceSlots[m.face_offset + f];
}
PRB_DEV V3 triInterp(V3 v0, V3 v1, V3 v2, float u, float v) { return (v1 * u + v2 * v) + v0 * (1 - u - v); }
PRB_DEV V3 quadInterp(V3 v0, V3 v1, V3 v2, V3 v3, float u, float v)
{
	return ((v0 * (1 - u) * (1 - v) + v1 * u * (1 - v)) + v2 * (1 - u) * v) + v3 * u * v;
}
PRB_DEV V3 faceInterpV(const FaceData& f, const V3* a, float u, float v) { return f.quad ? quadInterp(a[0], a[1], a[2], a[3], u, v) : triInterp(a[0], a[1], a[2], u, v); }
PRB_DEV float faceArea(const FaceData& f)
{
	if (f.quad)
		return 0.5f * sqrtf(norm2(cross(f.V[2] - f.V[0], f.V[3] - f.V[1])));
	return 0.5f * sqrtf(norm2(cross(f.V[1] - f.V[0], f.V[2] - f.V[0])));
}

PRB_DEV void provideGeometryPointBody(const DScene& S, uint32_t entityID, uint32_t prim, float qu, float qv, V3 position, GeomPoint& pt)
{
	const prb_entity& en = S.entities[entityID];
	pt.entity			 = entityID;
	pt.emission			 = en.emission_id;
	if (en.type == PRB_ENTITY_MESH) { // mesh.cpp:205-250
		const prb_mesh m = S.meshes[en.mesh_id];
		FaceData f;
		getFace(S, m, prim, f, true, true);
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
					const V3 nd = dp1 * dv2 - dp2 * dv1;
				V3 nx		= mk(divPositive(nd.x, det), divPositive(nd.y, det), divPositive(nd.z, det)); // det > 0: see divPositive
					nx	  = nx - pt.N * dot(pt.N, nx);
					nx	  = normalized(nx);
					pt.Nx = nx;
					pt.Ny = cross(pt.N, nx);
				}
			} else {
				frame_duff(pt.N, pt.Nx, pt.Ny);
			}
		} else { // dPdu, dPdv of the vertex buffer (rtcInterpolate, mesh.cpp:55-80)
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
		pt.material = f.slot < en.material_count ? S.entityMaterials[en.material_offset + f.slot] : PRB_INVALID_ID;
		pt.N		= normalized(m3mul(en.normal_matrix, pt.N));
		pt.Nx		= normalized(m3mul(en.normal_matrix, pt.Nx));
		pt.Ny		= normalized(m3mul(en.normal_matrix, pt.Ny));
		pt.prim		= prim;
	} else if (en.type == PRB_ENTITY_SPHERE) { // sphere.cpp:128-143
		pt.N = normalized(position - xfPoint(en.local_to_world, mk(0, 0, 0)));
		tangent_frame(pt.N, pt.Nx, pt.Ny);
		uv_from_normal(pt.N, pt.u, pt.v);
		pt.prim		= 0;
		pt.material = S.entityMaterials[en.material_offset];
	} else { // plane.cpp:206-220
		pt.N		= ld3(en.geo + 9);
		pt.Nx		= ld3(en.geo + 3);
		pt.Ny		= ld3(en.geo + 6);
		pt.u		= qu;
		pt.v		= qv;
		pt.prim		= 0;
		pt.material = S.entityMaterials[en.material_offset];
	}
}
// Out-of-line form for the generic light sampling (sampleLight); the shading kernels call the body in line: out of line it reads
// every DScene field it needs through a generic pointer to the kernel parameter (LD.E + a descriptor R2UR pair per load: 6.5 %
// of k_shade's instructions on the Cornell box, and a dependent load in front of every data access; C2 328 -> 354 Msamples/s)
__device__ __noinline__ void provideGeometryPoint(const DScene& S, uint32_t entityID, uint32_t prim, float qu, float qv, V3 position, GeomPoint& pt)
{
	provideGeometryPointBody(S, entityID, prim, qu, qv, position, pt);
}

// ------------------------------------------------------------------ materials
constexpr float AIR = 1.0002926f; // dielectric.cpp:17
constexpr uint32_t MSF_Delta = 0x2, MSF_SpectralVarying = 0x4;
struct MatEval {
	Blob weight, pdf;
	uint32_t flags, type;
};
struct MatSample {
	V3 L;
	Blob weight, pdf;
	uint32_t flags, type;
	PRB_DEV bool isDelta() const { return flags & MSF_Delta; }
	PRB_DEV bool isHeroCollapsing() const { return (flags & MSF_Delta) && (flags & MSF_SpectralVarying); }
};
struct MatCtx {
	V3 V, L;
	Blob wvl;
	float u, v;
	uint32_t rayFlags;
	// One-entry cache of a shading-node value: NEE (IMaterial::eval) and scattering (IMaterial::sample) of a path vertex
	// evaluate the same albedo node at the same wavelengths and surface parameters; the second evaluation (an IEEE sqrt and
	// division per wavelength for an upsampled RGB colour) is a copy of the first.  Valid for one vertex (wvl, u, v fixed).
	mutable uint32_t cachedNode = PRB_INVALID_ID;
	mutable Blob cachedValue;
	// set by the kernels that inline the Lambert code (the host only picks them for scenes without image nodes): node
	// evaluation goes straight to the leaf-only evalNodeBase; a compile-time constant there, so the test folds away
	bool noImageNodes = false;
};
template <bool FAST = false>
PRB_DEV Blob evalNodeCached(const DScene& S, const MatCtx& c, uint32_t node);
PRB_DEV uint32_t contribFlags(const prb_material& m) { return (m.flags & PRB_MATF_SPECTRAL_VARYING) ? MSF_SpectralVarying : 0; }
PRB_DEV void rejectSample(MatSample& s, uint32_t type, uint32_t flags)
{
	s.L		 = mk(0, 0, 0);
	s.weight = blob(0);
	s.pdf	 = blob(0);
	s.type	 = type;
	s.flags	 = flags;
}
PRB_DEV RoughDistribution roughOf(const prb_material& m)
{
	RoughDistribution r;
	r.M1	= m.f[0];
	r.M2	= m.f[1];
	r.aniso = m.flags & PRB_MATF_ANISOTROPIC;
	r.vndf	= m.flags & PRB_MATF_VNDF;
	return r;
}

// --- principled closure (principled.cpp:34-447)
struct Principled {
	Blob Base, IOR;
	float DiffuseTransmission, Roughness, Anisotropic, SpecularTransmission, SpecularTint, Flatness, Metallic, Sheen, SheenTint, Clearcoat, ClearcoatGloss;
	bool vndf, thin, hasTrans;
	static constexpr float EVAL_EPS = 1e-4f;
	PRB_DEV static float mixf(float v0, float v1, float t) { return (1 - t) * v0 + t * v1; }
	PRB_DEV static float schlickR0(float eta)
	{
		const float f = (eta - 1.0f) / (eta + 1.0f);
		return f * f;
	}
	PRB_DEV float thinTransmissionRoughness() const { return fmaxf(0.0f, fminf(1.0f, (0.65f * (bsum(IOR) / 4) - 0.35f) * Roughness)); }
	PRB_DEV RoughDistribution roughnessClosure(float r) const
	{
		const float aspect = sqrtf(1 - Anisotropic * 0.9f);
		RoughDistribution d;
		d.M1	= fmaxf(0.001f, r * r / aspect);
		d.M2	= fmaxf(0.001f, r * r * aspect);
		d.aniso = true;
		d.vndf	= vndf;
		return d;
	}
	PRB_DEV bool isDelta() const { return roughnessClosure(Roughness).isDelta(); }
	PRB_DEV void lobes(V3 V, float& dr, float& dt, float& sr, float& st) const
	{
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
	PRB_DEV Blob tintColor(const DScene& S, const Blob& wvl) const
	{
		float lum = 0;
#pragma unroll
		for (int i = 0; i < 4; ++i)
			lum = fmaxf(lum, Base[i] * cieEval(S, 1, wvl[i]));
		return lum > PR_EPSILON ? Base / lum : blob(1);
	}
	PRB_DEV Blob disneyFresnelTerm(const DScene& S, float HdotV, float HdotL, const Blob& wvl) const
	{
		Blob res;
		if (Metallic <= EVAL_EPS) {
#pragma unroll
			for (int i = 0; i < 4; ++i)
				res[i] = fresnel_dielectric(HdotV, AIR, IOR[i]);
			return res;
		}
		const Blob color = tintColor(S, wvl);
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const float eta = HdotV < 0 ? AIR / IOR[i] : fdiv(IOR[i], AIR);
			const float r0	= mixf(schlickR0(eta) * mixf(1.0f, color[i], SpecularTint), Base[i], Metallic);
			const float f1	= fresnel_dielectric(HdotV, AIR, IOR[i]);
			const float f2	= schlick(fabsf(HdotL), r0);
			res[i]			= mixf(f1, f2, Metallic);
		}
		return res;
	}
	PRB_DEV float retroDiffuseTerm(const MatCtx& c, float HdotL) const
	{
		const float alpha2 = Roughness * Roughness;
		const float fd90   = 0.5f + 2 * HdotL * HdotL * alpha2;
		const float lk = schlick_term(absCosTheta(c.L)), vk = schlick_term(absCosTheta(c.V));
		return PR_INV_PI * fd90 * (lk + vk + lk * vk * (fd90 - 1.0f));
	}
	PRB_DEV float subsurfaceTerm(const MatCtx& c, float HdotL) const
	{
		const float alpha2 = Roughness * Roughness;
		const float fss90  = HdotL * HdotL * alpha2;
		const float lk = schlick_term(absCosTheta(c.L)), vk = schlick_term(absCosTheta(c.V));
		const float fss = mixf(1.0f, fss90, lk) * mixf(1.0f, fss90, vk);
		const float f	= absCosTheta(c.L) + absCosTheta(c.V);
		if (fabsf(f) < PR_EPSILON)
			return 0.0f;
		return 1.25f * (fss * (1.0f / f - 0.5f) + 0.5f);
	}
	PRB_DEV float diffuseTerm(const MatCtx& c, float HdotL) const
	{
		const float lk = schlick_term(absCosTheta(c.L)), vk = schlick_term(absCosTheta(c.V));
		float diffuse = 1;
		if (thin)
			diffuse = mixf(1.0f, subsurfaceTerm(c, HdotL), Flatness);
		return PR_INV_PI * diffuse * (1 - 0.5f * lk) * (1 - 0.5f * vk);
	}
	PRB_DEV float clearcoatTerm(const MatCtx& c, V3 H) const
	{
		const float F0 = 0.04f, R = 0.25f;
		const float D  = ndf_ggx1(H, mixf(0.1f, 0.001f, ClearcoatGloss));
		const float hk = schlick_term(fabsf(dot(H, c.L)));
		const float F  = mixf(F0, 1.0f, hk);
		const float G  = g_1_smith_opt(absCosTheta(c.L), R) * g_1_smith_opt(absCosTheta(c.V), R);
		return R * D * F * G;
	}
	__device__ __noinline__ Blob eval(const DScene& S, const MatCtx& c) const
	{ // principled.cpp:274-344
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
				Blob sheen		  = blob(0);
				if (Sheen > EVAL_EPS) {
					const Blob tint = tintColor(S, c.wvl);
					Blob sheenColor;
#pragma unroll
					for (int i = 0; i < 4; ++i)
						sheenColor[i] = mixf(1.0f, tint[i], SheenTint);
					sheen = (sheenColor * Sheen) * schlick_term(fabsf(HdotL));
				}
				sheen = sheen * diffuseWeight;
				value = value + (Base * retro + sheen) * absCosTheta(c.L);
				const float diff = diffuseTerm(c, HdotL) * (thin ? 1 - DiffuseTransmission : diffuseWeight);
				value			 = value + Base * (diff * absCosTheta(c.L));
			}
			if (hasTrans && thin && isTransmission) {
				const float diff = diffuseTerm(c, HdotL) * DiffuseTransmission;
				value			 = value + Base * (diff * absCosTheta(c.L));
			}
		}
		{ // specularReflectionTerm
			MicrofacetReflection micro{ roughnessClosure(Roughness) };
			const float HdotV = dot(c.V, rH), HdotL2 = dot(c.L, rH);
			const Blob F = disneyFresnelTerm(S, HdotV, HdotL2, c.wvl);
			value		 = value + F * micro.eval(c.V, c.L);
		}
		if (hasTrans) {
			const float transmissionWeight = (1.0f - Metallic) * SpecularTransmission;
			if (transmissionWeight > EVAL_EPS) {
				Blob weight;
				const float scaledR = thin ? thinTransmissionRoughness() : Roughness;
				const RoughDistribution rd = roughnessClosure(scaledR);
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					MicrofacetTransmission micro{ rd, AIR, IOR[i] };
					const float R = micro.evalDielectric(c.V, c.L, false);
					weight[i]	  = thin ? sqrtf(Base[i]) * R : Base[i] * R;
				}
				if (c.rayFlags & PRB_RAY_LIGHT) {
#pragma unroll
					for (int i = 0; i < 4; ++i) {
						const float eta = HdotL < 0.0f ? fdiv(IOR[i], AIR) : AIR / IOR[i];
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
	__device__ __noinline__ Blob pdf(const MatCtx& c) const
	{ // principled.cpp:370-397
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
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					MicrofacetTransmission refr{ rd, AIR, IOR[i] };
					p[i] = refr.pdf(c.V, c.L);
				}
				pdfV = pdfV + p * st;
			}
		}
		return pdfV;
	}
	PRB_DEV V3 sampleDiffuse(Rng& rnd, V3 V) const
	{
		const bool flip = cosTheta(V) < 0;
		const float u2	= rnd.getFloat();
		const float u1	= rnd.getFloat();
		const V3 L		= cos_hemi(u1, u2);
		return flip ? -L : L;
	}
	PRB_DEV V3 sample(Rng& rnd, V3 V) const
	{ // principled.cpp:418-435
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
PRB_DEV void makePrincipled(const DScene& S, const prb_material& m, const MatCtx& c, Principled& p)
{
	p.Base				   = evalNode(S, m.node[0], c.wvl, c.u, c.v);
	p.IOR				   = evalNode(S, m.node[1], c.wvl, c.u, c.v);
	p.vndf				   = m.flags & PRB_MATF_VNDF;
	p.thin				   = m.flags & PRB_MATF_THIN;
	p.hasTrans			   = m.flags & PRB_MATF_HAS_TRANSMISSION;
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
}

// --- rough dielectric closure (roughdielectric.cpp:42-137)
struct RoughDielectric {
	RoughDistribution rd;
	Blob Spec, Trans, IOR;
	PRB_DEV Blob eval(V3 V, V3 L, bool isLightPath) const
	{
		Blob w;
		if (sameHemisphere(V, L)) {
			MicrofacetReflection refl{ rd };
#pragma unroll
			for (int i = 0; i < 4; ++i)
				w[i] = refl.evalDielectric(V, L, AIR, IOR[i]);
			return w * Spec;
		}
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			MicrofacetTransmission tr{ rd, AIR, IOR[i] };
			w[i] = tr.evalDielectric(V, L, isLightPath);
		}
		return w * Trans;
	}
	PRB_DEV Blob pdf(V3 V, V3 L) const
	{
		Blob F, p;
#pragma unroll
		for (int i = 0; i < 4; ++i)
			F[i] = fresnel_dielectric(cosTheta(V), AIR, IOR[i]);
		if (sameHemisphere(V, L)) {
			MicrofacetReflection refl{ rd };
			const float rp = refl.pdf(L, V);
			return F * blob(rp);
		}
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			MicrofacetTransmission tr{ rd, AIR, IOR[i] };
			p[i] = tr.pdf(V, L);
		}
		Blob omf;
#pragma unroll
		for (int i = 0; i < 4; ++i)
			omf[i] = 1 - F[i];
		return omf * p;
	}
	PRB_DEV V3 sample(Rng& rnd, V3 V) const
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

// LambertMaterial::eval / ::sample (lambert.cpp:33-43, :53-73); inline so that k_shade's all-Lambert instantiation can use
// them without the out-of-line material dispatch
#ifndef PRB_NODE_CACHE
#define PRB_NODE_CACHE 1
#endif
// FAST: called from a kernel that inlines the Lambert code (evalNodeFast)
template <bool FAST>
PRB_DEV Blob evalNodeCached(const DScene& S, const MatCtx& c, uint32_t node)
{
	if (!PRB_NODE_CACHE)
		return FAST ? evalNodeFast(S, node, c.wvl, c.u, c.v) : c.noImageNodes ? evalNodeBase(S, node, c.wvl, c.u, c.v) : evalNodeTex(S, node, c.wvl, c.u, c.v);
	if (c.cachedNode != node) {
		c.cachedValue = FAST ? evalNodeFast(S, node, c.wvl, c.u, c.v) : c.noImageNodes ? evalNodeBase(S, node, c.wvl, c.u, c.v) : evalNodeTex(S, node, c.wvl, c.u, c.v);
		c.cachedNode  = node;
	}
	return c.cachedValue;
}
template <bool FAST = false>
PRB_DEV void lambertEval(const DScene& S, const prb_material& m, const MatCtx& c, MatEval& out)
{
	const bool two = m.flags & PRB_MATF_TWO_SIDED;
	const float d  = sameHemisphere(c.V, c.L) ? (two ? fabsf(c.L.z) : fmaxf(0.0f, c.L.z)) : 0;
	out.weight	   = evalNodeCached<FAST>(S, c, m.node[0]) * d * PR_INV_PI;
	out.pdf		   = blob(cos_hemi_pdf(d));
}
template <bool FAST = false>
PRB_DEV void lambertSample(const DScene& S, const prb_material& m, const MatCtx& c, Rng& rnd, MatSample& out)
{
	if (!(m.flags & PRB_MATF_TWO_SIDED) && c.V.z < 0.0f) {
		rejectSample(out, 0, 0);
		return;
	}
	const float u2 = rnd.getFloat(); // cos_hemi(RND.getFloat(), RND.getFloat()): second argument drawn first
	const float u1 = rnd.getFloat();
	out.L		   = cos_hemi(u1, u2);
	out.weight	   = evalNodeCached<FAST>(S, c, m.node[0]);
	out.pdf		   = blob(cos_hemi_pdf(out.L.z));
	out.L		   = makeSameHemisphere(c.V, out.L);
}
// OrenNayarMaterial::calc (improved Oren-Nayar), orennayar.cpp:29-49
PRB_DEV Blob orenNayarCalc(const DScene& S, const prb_material& m, const MatCtx& c, V3 L, float NdotL)
{
	float roughness = m.f[0];
	roughness *= roughness;
	Blob weight = evalNodeCached(S, c, m.node[0]);
	if (roughness > PR_EPSILON) {
		const float s = -NdotL * c.V.z + dot(c.V, L);
		const float t = s < PR_EPSILON ? 1.0f : fmaxf(NdotL, c.V.z);
		const Blob A  = blob(1 - 0.5f * roughness / (roughness + 0.33f)) + ((weight * 0.17f) * roughness) / (roughness + 0.13f);
		const float B = 0.45f * roughness / (roughness + 0.09f);
		weight		  = weight * (A + blob(B * s / t));
	}
	return weight;
}

// FIXED >= 0: the caller knows the material type at compile time (the per-type stage kernels): the switch folds to one case
template <int FIXED>
__device__ __noinline__ void materialEvalLeafT(const DScene& S, uint32_t matID, const MatCtx& c, MatEval& out)
{
	const prb_material m = S.materials[matID];
	out.flags			 = 0;
	out.type			 = 0;
	switch (FIXED >= 0 ? (uint32_t)FIXED : m.type) {
	case PRB_MAT_DIFFUSE: lambertEval(S, m, c, out); break;
	case PRB_MAT_DIELECTRIC:
		out.pdf	   = blob(0);
		out.weight = blob(0);
		out.type   = 3;
		out.flags  = MSF_Delta | contribFlags(m);
		break;
	case PRB_MAT_MIRROR: // mirror.cpp:27-37
		out.pdf	   = blob(0);
		out.weight = blob(0);
		out.type   = 1;
		out.flags  = MSF_Delta;
		break;
	case PRB_MAT_ORENNAYAR: { // orennayar.cpp:51-60
		const float d = fmaxf(0.0f, c.L.z);
		out.weight	  = (orenNayarCalc(S, m, c, c.L, d) * PR_INV_PI) * d;
		out.pdf		  = blob(cos_hemi_pdf(d));
		break;
	}
	case PRB_MAT_CONDUCTOR:
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
		const Blob eta = evalNode(S, m.node[0], c.wvl, c.u, c.v), k = evalNode(S, m.node[1], c.wvl, c.u, c.v);
		Blob factor;
#pragma unroll
		for (int i = 0; i < 4; ++i)
			factor[i] = closure.evalConductor(c.L, c.V, eta[i], k[i]);
		out.weight = evalNode(S, m.node[2], c.wvl, c.u, c.v) * factor;
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
		cl.Spec	   = evalNode(S, m.node[0], c.wvl, c.u, c.v);
		cl.Trans   = (m.flags & PRB_MATF_TRANSMISSION_COLOR) ? evalNode(S, m.node[1], c.wvl, c.u, c.v) : cl.Spec;
		cl.IOR	   = evalNode(S, m.node[2], c.wvl, c.u, c.v);
		out.weight = cl.eval(c.V, c.L, c.rayFlags & PRB_RAY_LIGHT);
		out.pdf	   = cl.pdf(c.V, c.L);
		out.type   = sameHemisphere(c.V, c.L) ? 1 : 3;
		out.flags  = contribFlags(m);
		break;
	}
	case PRB_MAT_PRINCIPLED: { // principled.cpp:496-523
		Principled cl;
		makePrincipled(S, m, c, cl);
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
		out.weight = cl.eval(S, c);
		out.pdf	   = cl.pdf(c);
		break;
	}
	default:
		out.weight = blob(0);
		out.pdf	   = blob(0);
		break;
	}
}

template <int FIXED>
__device__ __noinline__ void materialSampleLeafT(const DScene& S, uint32_t matID, const MatCtx& c, Rng& rnd, MatSample& out)
{
	const prb_material m = S.materials[matID];
	out.flags			 = 0;
	out.type			 = 0;
	switch (FIXED >= 0 ? (uint32_t)FIXED : m.type) {
	case PRB_MAT_DIFFUSE: lambertSample(S, m, c, rnd, out); break;
	case PRB_MAT_DIELECTRIC: { // dielectric.cpp:60-114
		out.pdf		  = blob(1);
		const Blob n2 = evalNode(S, m.node[2], c.wvl, c.u, c.v);
		float F		  = fresnel_dielectric(cosTheta(c.V), AIR, n2[0]);
		const bool thin = m.flags & PRB_MATF_THIN;
		if (thin && F < 1.0f)
			F += (1 - F) * F / (F + 1);
		const Blob rWeight = evalNode(S, m.node[0], c.wvl, c.u, c.v);
		if (rnd.getFloat() <= F) {
			out.type   = 1;
			out.L	   = reflectZ(c.V);
			out.weight = rWeight;
		} else {
			Blob tWeight = (m.flags & PRB_MATF_TRANSMISSION_COLOR) ? evalNode(S, m.node[1], c.wvl, c.u, c.v) : rWeight;
			if (thin) {
				out.type   = 3;
				out.L	   = -c.V;
				out.weight = tWeight;
			} else {
				if (c.rayFlags & PRB_RAY_LIGHT) {
					const float eta = isPositiveHemisphere(c.V) ? AIR / n2[0] : fdiv(n2[0], AIR);
					tWeight			= tWeight * (eta * eta);
				}
				out.L = refractZ(AIR / n2[0], c.V);
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
		out.weight = evalNode(S, m.node[0], c.wvl, c.u, c.v);
		out.type   = 1;
		out.pdf	   = blob(1);
		out.L	   = reflectZ(c.V);
		out.flags  = MSF_Delta;
		break;
	case PRB_MAT_ORENNAYAR: { // orennayar.cpp:72-84
		const float u2 = rnd.getFloat(); // cos_hemi(RND.getFloat(), RND.getFloat()): second argument drawn first
		const float u1 = rnd.getFloat();
		out.L		   = cos_hemi(u1, u2);
		out.weight	   = orenNayarCalc(S, m, c, out.L, fmaxf(0.0f, out.L.z));
		out.pdf		   = blob(cos_hemi_pdf(out.L.z));
		break;
	}
	case PRB_MAT_CONDUCTOR: { // conductor.cpp:56-74
		const Blob eta = evalNode(S, m.node[0], c.wvl, c.u, c.v), k = evalNode(S, m.node[1], c.wvl, c.u, c.v);
		Blob fr;
#pragma unroll
		for (int i = 0; i < 4; ++i)
			fr[i] = fresnel_conductor(absCosTheta(c.V), 1, eta[i], k[i]);
		out.weight = fr * evalNode(S, m.node[2], c.wvl, c.u, c.v);
		out.type   = 1;
		out.pdf	   = blob(1);
		out.L	   = reflectZ(c.V);
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
			rejectSample(out, 1, out.flags);
			return;
		}
		const Blob eta = evalNode(S, m.node[0], c.wvl, c.u, c.v), k = evalNode(S, m.node[1], c.wvl, c.u, c.v);
		Blob factor;
#pragma unroll
		for (int i = 0; i < 4; ++i)
			factor[i] = closure.evalConductor(out.L, c.V, eta[i], k[i]);
		out.weight = evalNode(S, m.node[2], c.wvl, c.u, c.v) * factor;
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
		cl.Spec	  = evalNode(S, m.node[0], c.wvl, c.u, c.v);
		cl.Trans  = (m.flags & PRB_MATF_TRANSMISSION_COLOR) ? evalNode(S, m.node[1], c.wvl, c.u, c.v) : cl.Spec;
		cl.IOR	  = evalNode(S, m.node[2], c.wvl, c.u, c.v);
		out.L	  = cl.sample(rnd, c.V);
		out.flags = contribFlags(m);
		if (cl.rd.isDelta())
			out.flags |= MSF_Delta;
		if (out.L.x == 0 && out.L.y == 0 && out.L.z == 0) {
			rejectSample(out, 1, out.flags);
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
		Principled cl;
		makePrincipled(S, m, c, cl);
		out.L = cl.sample(rnd, c.V);
		if (cl.isDelta())
			out.flags |= MSF_Delta;
		if (out.L.x == 0 && out.L.y == 0 && out.L.z == 0) {
			rejectSample(out, 0, out.flags);
			return;
		}
		if (sameHemisphere(c.V, out.L))
			out.type = cl.Roughness < 0.5f ? 1 : 0;
		else
			out.type = cl.Roughness < 0.5f ? 3 : 2;
		MatCtx e   = c;
		e.L		   = out.L;
		out.weight = cl.eval(S, e);
		out.pdf	   = cl.pdf(e);
		if (out.pdf[0] > PR_EPSILON)
			out.weight = out.weight / out.pdf[0];
		if (cl.isDelta())
			out.pdf = blob(1);
		break;
	}
	default: rejectSample(out, 0, 0); break;
	}
}

// blend.cpp:20-148 / add.cpp:20-122 over two LEAF materials (node[0], node[1] hold their ids); everything else is a leaf.
// No recursion: the host rejects nested combinations, so the device stack stays statically sized.
PRB_DEV bool isCombination(uint32_t type) { return type == PRB_MAT_BLEND || type == PRB_MAT_ADD; }
PRB_DEV void materialEvalLeaf(const DScene& S, uint32_t matID, const MatCtx& c, MatEval& out) { materialEvalLeafT<-1>(S, matID, c, out); }
PRB_DEV void materialSampleLeaf(const DScene& S, uint32_t matID, const MatCtx& c, Rng& rnd, MatSample& out) { materialSampleLeafT<-1>(S, matID, c, rnd, out); }
// Children of a combination may be combinations themselves (blend.cpp / add.cpp take any IMaterial).  The device has no
// recursion budget, so the nesting is unrolled at compile time: a combination at depth DEPTH evaluates combination children
// with DEPTH - 1; the host refuses materials nested deeper than COMBINE_MAX_DEPTH levels.
constexpr int COMBINE_MAX_DEPTH = 3;
template <int DEPTH>
__device__ __noinline__ void materialEvalCombinedT(const DScene& S, uint32_t matID, const MatCtx& c, MatEval& out);
template <int DEPTH>
PRB_DEV void materialEvalChild(const DScene& S, uint32_t id, const MatCtx& c, MatEval& out)
{
	if constexpr (DEPTH > 1) {
		if (isCombination(S.materials[id].type)) {
			materialEvalCombinedT<DEPTH - 1>(S, id, c, out);
			return;
		}
	}
	materialEvalLeaf(S, id, c, out);
}
template <int DEPTH>
__device__ __noinline__ void materialEvalCombinedT(const DScene& S, uint32_t matID, const MatCtx& c, MatEval& out)
{ // out of line: its two child results only occupy stack while a combination is evaluated
	const prb_material& m = S.materials[matID];
	const uint32_t type	  = m.type;
	const bool add		  = type == PRB_MAT_ADD;
	const bool d0 = m.flags & PRB_MATF_CHILD0_DELTA, d1 = m.flags & PRB_MATF_CHILD1_DELTA;
	const float prob = fminf(1.0f, fmaxf(0.0f, m.f[0]));
	if (d0 && d1) { // MaterialDelta::All: never evaluated by the integrator
		out.pdf	   = blob(0);
		out.weight = blob(0);
		out.type   = 3;
		out.flags  = 0;
	} else if (d0 || d1) { // the non-delta child alone, scaled by its share
		materialEvalChild<DEPTH>(S, m.node[d0 ? 1 : 0], c, out);
		const float share = add ? 0.5f : (d0 ? prob : 1 - prob);
		out.pdf			  = out.pdf * share;
		if (!add)
			out.weight = out.weight * share;
	} else {
		MatEval o1, o2;
		materialEvalChild<DEPTH>(S, m.node[0], c, o1);
		materialEvalChild<DEPTH>(S, m.node[1], c, o2);
		out.flags = 0;
		if (add) {
			out.pdf	   = (o1.pdf + o2.pdf) * 0.5f; // (a + b) / 2
			out.weight = o1.weight + o2.weight;
			out.type   = o1.type;
		} else {
			out.pdf	   = o1.pdf * (1 - prob) + o2.pdf * prob;
			out.weight = o1.weight * (1 - prob) + o2.weight * prob;
			out.type   = prob <= 0.5f ? o1.type : o2.type;
		}
	}
}
PRB_DEV void materialEvalCombined(const DScene& S, uint32_t matID, const MatCtx& c, MatEval& out) { materialEvalCombinedT<COMBINE_MAX_DEPTH>(S, matID, c, out); }
// k_shade is instantiated per KIND: without the combination path for scenes that have no blend / add material (merely having
// the call in the kernel cost 4-5 % of k_shade on C2 / C4), and with the Lambert code inline for all-Lambert scenes.
// SHADE_MATERIALS_TYPE + t: every slot the kernel sees has a material of type t (the per-type queues of the staged path)
enum { SHADE_MATERIALS_LEAF = 0, SHADE_MATERIALS_COMBINED = 1, SHADE_MATERIALS_LAMBERT = 2, SHADE_MATERIALS_TYPE = 16 };
template <int KIND>
PRB_DEV void materialEval(const DScene& S, uint32_t matID, const MatCtx& c, MatEval& out)
{
	if (KIND >= SHADE_MATERIALS_TYPE) {
		materialEvalLeafT<KIND - SHADE_MATERIALS_TYPE>(S, matID, c, out);
	} else if (KIND == SHADE_MATERIALS_LAMBERT) { // every material of the scene is a Lambert material (Cornell box)
		out.flags = 0;
		out.type  = 0;
		lambertEval<true>(S, S.materials[matID], c, out);
	} else if (KIND == SHADE_MATERIALS_COMBINED && isCombination(S.materials[matID].type)) {
		materialEvalCombined(S, matID, c, out);
	} else {
		materialEvalLeaf(S, matID, c, out);
	}
}
__device__ __noinline__ void materialSampleCombined(const DScene& S, uint32_t matID, const MatCtx& c, Rng& rnd, MatSample& out)
{ // BlendMaterial / AddMaterial::sample (blend.cpp:112-130, add.cpp:97-113): one random number picks the child at every level;
  // the shares are applied innermost first, like the nested calls of the reference return
	float shares[COMBINE_MAX_DEPTH];
	bool adds[COMBINE_MAX_DEPTH];
	int depth	= 0;
	uint32_t id = matID;
	while (depth < COMBINE_MAX_DEPTH && isCombination(S.materials[id].type)) {
		const prb_material& m = S.materials[id];
		const bool add		  = m.type == PRB_MAT_ADD;
		const float prob	  = add ? 0.5f : fminf(1.0f, fmaxf(0.0f, m.f[0]));
		const bool first	  = rnd.getFloat() < (add ? 0.5f : 1 - prob);
		shares[depth]		  = add ? 0.5f : (first ? 1 - prob : prob);
		adds[depth]			  = add;
		++depth;
		id = m.node[first ? 0 : 1];
	}
	materialSampleLeaf(S, id, c, rnd, out);
	for (int i = depth - 1; i >= 0; --i) {
		if (!adds[i])
			out.weight = out.weight * shares[i];
		out.pdf = out.pdf * shares[i];
	}
}

template <int KIND>
PRB_DEV void materialSample(const DScene& S, uint32_t matID, const MatCtx& c, Rng& rnd, MatSample& out)
{
	if (KIND >= SHADE_MATERIALS_TYPE) {
		materialSampleLeafT<KIND - SHADE_MATERIALS_TYPE>(S, matID, c, rnd, out);
	} else if (KIND == SHADE_MATERIALS_LAMBERT) {
		out.flags = 0;
		out.type  = 0;
		lambertSample<true>(S, S.materials[matID], c, rnd, out);
	} else if (KIND == SHADE_MATERIALS_COMBINED && isCombination(S.materials[matID].type)) {
		materialSampleCombined(S, matID, c, rnd, out);
	} else {
		materialSampleLeaf(S, matID, c, rnd, out);
	}
}

// ------------------------------------------------------------------ samplers / mapper / camera
PRB_DEV uint32_t mjPermute(uint32_t i, uint32_t l, uint32_t p)
{ // Kensler CMJ permute, MultiJitteredSampler.cpp:21-76
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
PRB_DEV float haltonValue(uint32_t index, uint32_t base)
{ // HaltonSampler.cpp:14-25
	float result = 0;
	float f		 = 1;
	for (uint32_t i = index; i > 0;) {
		f = f / (float)base;
		result += f * (float)(i % base);
		i = (uint32_t)floorf((float)i / (float)base);
	}
	return result;
}
PRB_DEV void sampler2D(const DScene& S, const prb_sampler& s, Rng& rnd, uint32_t index, float& x, float& y)
{
	switch (s.type) {
	case PRB_SAMPLER_SOBOL: // SobolSampler.cpp:67-73
		if (s.max_samples <= index) {
			rnd.get2D(x, y);
		} else {
			const float* t = S.pool + s.table_offset + s.max_samples;
			x			   = __ldg(t + 2 * index);
			y			   = __ldg(t + 2 * index + 1);
		}
		break;
	case PRB_SAMPLER_MJITT: { // MultiJitteredSampler.cpp:118-150
		constexpr uint32_t FH = 0x51633e2d, F1 = 0x68bc21eb, F2 = 0x02e5be93;
		const uint32_t maxS = max(1u, s.max_samples);
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
			const float* t = S.pool + s.table_offset + s.max_samples;
			x			   = __ldg(t + 2 * index);
			y			   = __ldg(t + 2 * index + 1);
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
PRB_DEV float sampler1D(const DScene& S, const prb_sampler& s, Rng& rnd, uint32_t index)
{
	switch (s.type) {
	case PRB_SAMPLER_SOBOL:
		if (s.max_samples <= index)
			return rnd.getFloat();
		return __ldg(S.pool + s.table_offset + index);
	case PRB_SAMPLER_MJITT: {
		const float j = rnd.getFloat();
		return (index % s.bins_1d + j) / s.bins_1d;
	}
	case PRB_SAMPLER_HALTON:
		return index < s.max_samples ? __ldg(S.pool + s.table_offset + index) : haltonValue(index + s.seed, s.m2d_x);
	case PRB_SAMPLER_STRATIFIED: { // StratifiedSampler.cpp:22-26
		const float range = 1.0f / (int)s.bins_1d;
		return rnd.getFloat() * range + (int)index * range;
	}
	case PRB_SAMPLER_UNIFORM: return 0.5f;
	default: return rnd.getFloat();
	}
}
PRB_DEV int cdfSearch(const float* cdf, int size, float u)
{ // Interval::binary_search, src/base/container/Interval.h
	int first = 0, len = size;
	while (len > 0) {
		const int half = len / 2, middle = first + half;
		if (__ldg(cdf + middle) <= u) {
			first = middle + 1;
			len -= half + 1;
		} else {
			len = half;
		}
	}
	return max(0, min(first - 1, size - 2));
}
PRB_DEV float sampleContinuous(const float* cdf, int size, float u, float& pdf)
{ // Distribution1D::sampleContinuous, Distribution1D.inl:76-86,119-135
	const int off = cdfSearch(cdf, size, u);
	const float c0 = __ldg(cdf + off), c1 = __ldg(cdf + off + 1);
	float rem	  = u - c0;
	const float k = c1 - c0;
	if (k > PR_EPSILON)
		rem /= k;
	pdf = c1 - c0;
	pdf *= (size - 1);
	return (off + rem) / (size - 1);
}

struct CameraSampleOut {
	V3 origin, dir;
	float tmin, tmax;
	Blob wvl, wvlPDF;
	bool mono;
};
PRB_DEV void constructCameraRay(const DScene& S, uint32_t px, uint32_t py, uint32_t iteration, Rng& rnd, CameraSampleOut& o)
{ // RenderTile::constructCameraRay, RenderTile.cpp:71-132 + PerspectiveCamera::constructRay, perspective.cpp:45-82
	const prb_settings& st = S.settings;
	float ax, ay, lx, ly;
	sampler2D(S, S.aa, rnd, iteration, ax, ay);
	const float pixx = ((float)px + ax) - 0.5f, pixy = ((float)py + ay) - 0.5f;
	sampler2D(S, S.lens, rnd, iteration, lx, ly);
	(void)sampler1D(S, S.time, rnd, iteration); // time sample is drawn; static scenes do not use it
	if (st.spectral_mono) {
		o.wvl	 = blob(st.spectral_start);
		o.wvlPDF = blob(1.0f);
	} else {
		const float start = st.spectral_start, end = st.spectral_end;
		switch (S.mapper.type) {
		case PRB_MAPPER_SPD_CMIS: // spd.cpp:40-47
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				float pdf;
				const float x = sampleContinuous(S.pool + S.mapper.cdf_offset, (int)S.mapper.cdf_size, rnd.getFloat(), pdf);
				o.wvl[i]	  = x * (end - start) + start;
				o.wvlPDF[i]	  = pdf;
			}
			break;
		case PRB_MAPPER_CIE: // cie.cpp:24-31,61-69 + CIE::sample_trunc, CIE.h:117-127 (0..1 for the full range)
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				float pdf;
				const float cs = S.mapper.trunc_cdf_start, ce = S.mapper.trunc_cdf_end;
				const float v  = sampleContinuous(S.pool + S.mapper.cdf_offset, (int)S.mapper.cdf_size, cs + rnd.getFloat() * (ce - cs), pdf);
				pdf /= (ce - cs);
				o.wvl[i]	= v * (end - start) + start;
				o.wvlPDF[i] = pdf;
			}
			break;
		case PRB_MAPPER_AGH_CMIS: // agh.cpp:49-57: aghSample / aghPDF per wavelength
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				const float C = S.mapper.trunc_cdf_start, N = S.mapper.trunc_cdf_end;
				o.wvl[i]	  = 538.0f - fdiv(cr_atanh(C - N * rnd.getFloat()), 0.0072f);
				const float K = cr_cosh(0.0072f * (o.wvl[i] - 538.0f));
				o.wvlPDF[i]	  = 1 / (K * K * N);
			}
			break;
		case PRB_MAPPER_AGH_HERO: { // agh.cpp:98-103 + Standard.h:8-21
			const float C = S.mapper.trunc_cdf_start, N = S.mapper.trunc_cdf_end;
			const float hero = 538.0f - fdiv(cr_atanh(C - N * rnd.getFloat()), 0.0072f);
			const float K	 = cr_cosh(0.0072f * (hero - 538.0f));
			const float span = end - start, delta = span / 4, s = hero - start;
			o.wvl[0] = hero;
#pragma unroll
			for (int i = 1; i < 4; ++i)
				o.wvl[i] = start + fmodf(s + i * delta, span);
			o.wvlPDF = blob(1 / (K * K * N));
			break;
		}
		case PRB_MAPPER_SPD_HERO: { // spd.cpp:104-112 + Standard.h:8-21
			float pdf;
			const float u	 = rnd.getFloat();
			const float hero = sampleContinuous(S.pool + S.mapper.cdf_offset, (int)S.mapper.cdf_size, u, pdf) * (end - start) + start;
			const float span = end - start, delta = span / 4, s = hero - start;
			o.wvl[0] = hero;
#pragma unroll
			for (int i = 1; i < 4; ++i)
				o.wvl[i] = start + fmodf(s + i * delta, span);
			o.wvlPDF = blob(pdf);
			break;
		}
		default: { // random.cpp:22-36
			const float u	 = rnd.getFloat();
			const float span = end - start, delta = span / 4, s = u * span;
			o.wvl[0] = s + start;
#pragma unroll
			for (int i = 1; i < 4; ++i)
				o.wvl[i] = start + fmodf(s + i * delta, span);
			o.wvlPDF = blob(1.0f);
			break;
		}
		}
	}
	const float nx = 2 * (pixx / (float)st.film_width - 0.5f);
	const float ny = -(2 * (pixy / (float)st.film_height - 0.5f));
	if (S.camera.type == PRB_CAMERA_ORTHOGRAPHIC) { // OrthoCamera::constructRay, plugins/main/cameras/ortho.cpp:47-66
		o.origin = (ld3(S.camera.origin) + ld3(S.camera.right) * nx) + ld3(S.camera.up) * ny;
		o.dir	 = ld3(S.camera.dir);
	} else {
		V3 dir	 = (ld3(S.camera.right) * nx + ld3(S.camera.up) * ny) + ld3(S.camera.dir);
		o.origin = ld3(S.camera.origin);
		if (S.camera.has_dof) { // PerspectiveCamera<HasDOF = true>::constructRay, perspective.cpp:66-75
			float s, c;
			cr_sincos(2 * PR_PI * lx, &s, &c);
			const V3 e = (ld3(S.camera.aperture_x) * ly) * s + (ld3(S.camera.aperture_y) * ly) * c;
			o.origin   = o.origin + e;
			dir		   = dir - e;
		}
		o.dir = normalized(dir);
	}
	o.tmin		   = S.camera.near_t;
	o.tmax		   = S.camera.far_t;
	o.mono		   = st.spectral_mono || !st.spectral_hero;
}

// ------------------------------------------------------------------ lights
struct SQ { // spherical rectangle, plane.cpp:100-145
	V3 o, n;
	float z0, x0, y0, x1, y1, b0, b1, k, S;
};
PRB_DEV float safe_acos(float a) { return cr_acos(fmaxf(-1.0f, fminf(1.0f, a))); }
PRB_DEV void computeSQ(const prb_entity& en, V3 o, SQ& sq)
{
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
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const float diff = a[i] - b[i];
		nz[i]			 = c[i] * diff;
		nz[i] /= sqrtf(sq.z0 * sq.z0 * diff * diff + nz[i] * nz[i]);
	}
	const float g0 = safe_acos(-nz[0] * nz[1]), g1 = safe_acos(-nz[1] * nz[2]), g2 = safe_acos(-nz[2] * nz[3]), g3 = safe_acos(-nz[3] * nz[0]);
	sq.b0 = nz[0];
	sq.b1 = nz[2];
	sq.k  = 2 * PR_PI - g2 - g3;
	sq.S  = g0 + g1 - sq.k;
}
// IEntity::sampleParameterPointPDF(p, info): mesh 1/worldArea; sphere 2*pdfCache; plane spherical rectangle
PRB_DEV float entityPositionPDF(const DScene& S, uint32_t entityID, V3 p, V3 infoOrigin)
{
	const prb_entity& en = S.entities[entityID];
	if (en.type == PRB_ENTITY_SPHERE)
		return 2 * en.geo[5];
	if (en.type == PRB_ENTITY_PLANE) { // plane.cpp:184-195
		SQ sq;
		computeSQ(en, infoOrigin, sq);
		const float pdf_s = sq.S > PR_EPSILON ? 1 / sq.S : 0.0f;
		const V3 L		  = p - infoOrigin;
		const float dist2 = norm2(L);
		const float ndotv = fabsf(dot(normalized(L), ld3(en.geo + 26)));
		return ndotv <= PR_EPSILON ? 0 : pdf_s * fabsf(ndotv) / dist2;
	}
	return en.pdf_area;
}
struct LightSample {
	Blob radiance;
	V3 outgoing, lightPos;
	float posPDF, dirPDF_S, cosLight;
	bool infinite, delta;
};
// ---- sky / sun (plugins/main/infinitelights/sky.cpp, sun.cpp; tables precomputed by the host, see prb200_abi.h)
constexpr float SKY_ELEVATION_RANGE = PR_PI * 0.5f; // skysun/ElevationAzimuth.h:6-7
constexpr float SKY_AZIMUTH_RANGE	= PR_PI * 2;
PRB_DEV float skyModelRadiance(const DScene& S, const prb_light& l, int band, float el, float az)
{ // SkyModel::radiance, skysun/SkyModel.h:19-24
	const int azc = (int)l.az_count, elc = (int)l.el_count;
	const int az_in = max(0, min(azc - 1, (int)(fdiv(az, SKY_AZIMUTH_RANGE) * (float)azc)));
	const int el_in = max(0, min(elc - 1, (int)(fdiv(el, SKY_ELEVATION_RANGE) * (float)elc)));
	return __ldg(S.pool + l.table_offset + ((size_t)el_in * azc + az_in) * PRB_SKY_BANDS + band);
}
PRB_DEV Blob skyRadiance(const DScene& S, const prb_light& l, const Blob& wvls, float el, float az)
{ // SkyLight::radiance, sky.cpp:168-184
	Blob b;
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const float af	= fmaxf(0.0f, fdiv(wvls[i] - PRB_SKY_BAND_START, PRB_SKY_BAND_DELTA));
		const int index = (int)fminf((float)(PRB_SKY_BANDS - 2), af);
		const float t	= fminf((float)(PRB_SKY_BANDS - 1), af) - index;
		b[i]			= skyModelRadiance(S, l, index, el, az) * (1 - t) + skyModelRadiance(S, l, index + 1, el, az) * t;
	}
	return b;
}
// Distribution2D::sampleContinuous / continuousPdf, core/sampler/Distribution2D.cpp:13-31
PRB_DEV void dist2DSampleContinuous(const DScene& S, const prb_light& l, float u0, float u1, float& d0, float& d1, float& pdf)
{
	const float* marginal = S.pool + l.dist_offset;
	const int h = (int)l.dist_h, w = (int)l.dist_w;
	float pdf1, pdf0;
	d1						 = sampleContinuous(marginal, h + 1, u1, pdf1);
	const int moff			 = cdfSearch(marginal, h + 1, u1);
	const float* conditional = marginal + (h + 1) + (size_t)moff * (w + 1);
	d0						 = sampleContinuous(conditional, w + 1, u0, pdf0);
	pdf						 = pdf0 * pdf1;
}
PRB_DEV float dist2DContinuousPdf(const DScene& S, const prb_light& l, float x0, float x1)
{ // Distribution1D::continuousPdf, Distribution1D.inl:93-99
	const float* marginal = S.pool + l.dist_offset;
	const uint32_t h = l.dist_h, w = l.dist_w;
	const uint32_t moff		 = min(h - 1, (uint32_t)(x1 * (float)h));
	const float pdf1		 = (__ldg(marginal + moff + 1) - __ldg(marginal + moff)) * (float)h;
	const float* conditional = marginal + (h + 1) + (size_t)moff * (w + 1);
	const uint32_t off		 = min(w - 1, (uint32_t)(x0 * (float)w));
	const float pdf0		 = (__ldg(conditional + off + 1) - __ldg(conditional + off)) * (float)w;
	return pdf0 * pdf1;
}
PRB_DEV bool isInfLight(const prb_light& l) { return l.type != PRB_LIGHT_AREA; }
PRB_DEV bool isDeltaLight(const prb_light& l) { return l.type == PRB_LIGHT_SUN_DELTA; }

// Light::sample with SamplingInfo + Point (NEE), src/core/light/Light.cpp:108-226
// EnvironmentLight<UseDistribution = true>::sampleDir / samplePosDir (environment.cpp:83-104): (u, v) from the Distribution2D of
// the image, direction = Spherical::cartesian_from_uv.  Out of line: the inline light sampling below is part of the
// all-Lambert k_shade, whose executed footprint bounds it (instruction fetch).
__device__ __noinline__ void sampleEnvironmentMap(const DScene& S, const prb_light& l, V3 P, const Blob& wvl, Rng& rnd, LightSample& o)
{
	float dx, dy, px, py;
	rnd.get2D(dx, dy);
	rnd.get2D(px, py);
	float u0, u1, pdf;
	dist2DSampleContinuous(S, l, dx, dy, u0, u1, pdf);
	const V3 local		 = cartesian_from_uv(u0, u1);
	const float sinTheta = cr_sin(u1 * PR_PI);
	const float denom	 = 2 * PR_PI * PR_PI * sinTheta;
	o.delta				 = false;
	o.dirPDF_S			 = pdf * ((denom <= PR_EPSILON) ? 0.0f : 1.0f / denom);
	o.outgoing			 = m3mul(l.normal_matrix, local);
	o.radiance			 = evalNode(S, l.radiance_node, wvl, u0, u1); // coord.UV = uv
	o.lightPos			 = P + l.scene_radius * o.outgoing;
	o.posPDF			 = 1;
	o.cosLight			 = 1;
	o.infinite			 = true;
}
// ENVMAP: with the image-based environment light and image-capable node evaluation (the generic, out-of-line sampleLight);
// without, the form the all-Lambert kernels inline -- the host only picks those for scenes without image nodes, and every
// out-of-line call in their body costs registers there whether it is taken or not
template <bool ENVMAP>
PRB_DEV void sampleLightBody(const DScene& S, const prb_light& l, V3 P, const Blob& wvl, Rng& rnd, LightSample& o)
{
	o.delta = false;
	if (l.type == PRB_LIGHT_SKY) { // SkyLight::sampleDir / samplePosDir, sky.cpp:82-113
		float dx, dy, px, py;
		rnd.get2D(dx, dy);
		rnd.get2D(px, py);
		float pdf, u0, u1;
		dist2DSampleContinuous(S, l, dx, dy, u0, u1, pdf);
		const float el	  = l.sky_extend ? 2 * SKY_ELEVATION_RANGE * (u1 - 0.5f) : SKY_ELEVATION_RANGE * u1;
		const float az	  = SKY_AZIMUTH_RANGE * u0;
		const float theta = 0.5f * PR_PI - el; // ElevationAzimuth::toDirection
		float st, ct, sp, cp;
		cr_sincos(theta, &st, &ct);
		cr_sincos(az, &sp, &cp);
		o.outgoing		  = m3mul(l.normal_matrix, spherical_cartesian(st, ct, sp, cp));
		const float f	  = cr_cos(el);
		const float denom = 2 * PR_PI * PR_PI * f;
		o.dirPDF_S		  = pdf * ((denom <= PR_EPSILON) ? 0.0f : 1.0f / denom);
		o.radiance		  = skyRadiance(S, l, wvl, el, az);
		o.lightPos		  = P + l.scene_radius * o.outgoing;
		o.posPDF		  = 1;
		o.cosLight		  = 1;
		o.infinite		  = true;
		return;
	}
	if (l.type == PRB_LIGHT_SUN || l.type == PRB_LIGHT_SUN_DELTA) { // SunLight / SunDeltaLight::sampleDir, sun.cpp:77-99,177-199
		float dx, dy, px, py;
		rnd.get2D(dx, dy);
		rnd.get2D(px, py);
		const V3 sunDir = ld3(l.sun_dir);
		if (l.type == PRB_LIGHT_SUN) { // Sampling::uniform_cone, src/base/math/Sampling.h:101-107
			const float cosTheta = fmaf(dx, l.sun_cos_theta, 1 - dx);
			const float sinTheta = sqrtf(fmaxf(0.0f, diffProd(1, 1, cosTheta, cosTheta)));
			float sp, cp;
			cr_sincos(2 * PR_PI * dy, &sp, &cp);
			o.outgoing = fromTangentSpace(sunDir, ld3(l.sun_dx), ld3(l.sun_dy), mk(cp * sinTheta, sp * sinTheta, cosTheta));
			o.dirPDF_S = l.sun_pdf;
		} else {
			o.outgoing = sunDir;
			o.dirPDF_S = 1;
			o.delta	   = true;
		}
#pragma unroll
		for (int i = 0; i < 4; ++i)
			o.radiance[i] = tableLookup(S.pool + l.table_offset, l.table_count, l.table_start, l.table_end, wvl[i]);
		o.lightPos = P + l.scene_radius * o.outgoing;
		o.posPDF   = 0;
		o.cosLight = 1;
		o.infinite = true;
		return;
	}
	if (l.type == PRB_LIGHT_ENV) { // environment.cpp sampleDir / samplePosDir, :83-104
		if (ENVMAP && l.dist_w) { // UseDistribution (image based radiance)
			sampleEnvironmentMap(S, l, P, wvl, rnd, o);
			return;
		}
		float dx, dy, px, py;
		rnd.get2D(dx, dy);
		rnd.get2D(px, py);
		const V3 local = cos_hemi(dx, dy);
		o.dirPDF_S	   = cos_hemi_pdf(local.z);
		o.outgoing	   = m3mul(l.normal_matrix, local);
		o.radiance	   = ENVMAP ? evalNodeTex(S, l.radiance_node, wvl, dx, dy) : evalNodeFast(S, l.radiance_node, wvl, dx, dy);
		o.lightPos	   = P + l.scene_radius * o.outgoing;
		o.posPDF	   = 1;
		o.cosLight	   = 1;
		o.infinite	   = true;
		return;
	}
	o.infinite			 = false;
	const prb_entity& en = S.entities[l.entity_id];
	float rx, ry;
	rnd.get2D(rx, ry);
	V3 pos;
	float su, sv, pdfA;
	uint32_t prim = 0;
	if (en.type == PRB_ENTITY_MESH) { // mesh.cpp:187-203
		const prb_mesh m = S.meshes[en.mesh_id];
		float k1, k2;
		const float f1 = modff(rx * m.face_count, &k1); // SplitSample1D, SplitSample.h:6-26
		const float f2 = modff(ry * m.face_count, &k2);
		(void)k2;
		const uint32_t faceID = min((uint32_t)k1, m.face_count - 1);
		FaceData f;
		getFace(S, m, faceID, f, false, false);
		pdfA = 1.0f / (m.face_count * faceArea(f) * en.jacobian_det);
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
		const V3 local_o = normalized(xfPoint(en.world_to_local, P));
		if (dot(local_o, n) < -PR_EPSILON)
			n = -n;
		pos = xfPoint(en.local_to_world, en.geo[4] * n);
		uv_from_normal(n, su, sv);
		pdfA = 2 * en.geo[5];
	} else { // plane.cpp:147-182
		SQ sq;
		computeSQ(en, P, sq);
		const V3 mEx = ld3(en.geo + 3), mEy = ld3(en.geo + 6);
		const float au = fmaf(rx, sq.S, sq.k);
		const float fu = fmaf(cr_cos(au), sq.b0, -sq.b1) / cr_sin(au);
		const float cu = fminf(1.0f, fmaxf(-1.0f, copysignf(1.0f, fu) / sqrtf(sumProd(fu, fu, sq.b0, sq.b0))));
		const float xu = fminf(sq.x1, fmaxf(sq.x0, -(cu * sq.z0) / fmaxf(1e-7f, sqrtf(fmaf(-cu, cu, 1.0f)))));
		const float dd = sqrtf(sumProd(xu, xu, sq.z0, sq.z0));
		const float h0 = sq.y0 / sqrtf(sumProd(dd, dd, sq.y0, sq.y0));
		const float h1 = sq.y1 / sqrtf(sumProd(dd, dd, sq.y1, sq.y1));
		const float hv = fmaf(ry, h1 - h0, h0);
		const float hv2 = hv * hv;
		const float yv	= (hv2 < 1.0f - 1e-6f) ? (hv * dd) / sqrtf(1.0f - hv2) : sq.y1;
		pos				= ((sq.o + xu * mEx) + yv * mEy) + sq.z0 * sq.n;
		const float pdf_s = sq.S > PR_EPSILON ? 1 / sq.S : 0.0f;
		const V3 L		  = pos - P;
		const float dist2 = norm2(L);
		const float ndotv = fabsf(dot(normalized(L), ld3(en.geo + 26)));
		pdfA			  = ndotv <= PR_EPSILON ? 0 : pdf_s * ndotv / dist2;
		const V3 lp		  = xfPoint(en.world_to_local, pos) - ld3(en.geo + 29); // Plane::project
		su				  = dot(ld3(en.geo + 32), lp) * en.geo[38];
		sv				  = dot(ld3(en.geo + 35), lp) * en.geo[39];
	}
	GeomPoint gp;
	if (ENVMAP)
		provideGeometryPoint(S, l.entity_id, prim, su, sv, pos, gp);
	else
		provideGeometryPointBody(S, l.entity_id, prim, su, sv, pos, gp);
	o.outgoing = normalized(pos - P);
	o.dirPDF_S = 1;
	o.cosLight = fminf(1.0f, fmaxf(-1.0f, -dot(o.outgoing, gp.N)));
	o.radiance = ENVMAP ? evalNodeTex(S, S.emissions[l.emission_id].radiance_node, wvl, gp.u, gp.v) : evalNodeFast(S, S.emissions[l.emission_id].radiance_node, wvl, gp.u, gp.v);
	o.posPDF   = pdfA;
	o.lightPos = pos;
}
PRB_DEV void sampleLightInline(const DScene& S, const prb_light& l, V3 P, const Blob& wvl, Rng& rnd, LightSample& o) { sampleLightBody<false>(S, l, P, wvl, rnd, o); }
__device__ __noinline__ void sampleLight(const DScene& S, const prb_light& l, V3 P, const Blob& wvl, Rng& rnd, LightSample& o) { sampleLightBody<true>(S, l, P, wvl, rnd, o); }
PRB_DEV void envEval(const DScene& S, const prb_light& l, V3 dir, uint32_t depth, const Blob& wvl, Blob& rad, float& pdfS)
{ // EnvironmentLight::eval (no distribution), environment.cpp
	const V3 ld = m3mul(l.inv_normal_matrix, dir);
	float u, v;
	uv_from_normal(ld, u, v);
	const uint32_t node = (l.env_split && depth == 0) ? l.background_node : l.radiance_node;
	rad					= evalNode(S, node, wvl, u, v);
	if (l.dist_w) { // UseDistribution, environment.cpp:68-72
		pdfS				 = dist2DContinuousPdf(S, l, u, v);
		const float sinTheta = cr_sin(v * PR_PI);
		const float denom	 = 2 * PR_PI * PR_PI * sinTheta;
		pdfS *= (denom <= PR_EPSILON) ? 0.0f : 1.0f / denom;
	} else {
		pdfS = cos_hemi_pdf(fabsf(ld.z));
	}
}
// IInfiniteLight::eval for every infinite light type
__device__ __noinline__ void infLightEval(const DScene& S, const prb_light& l, V3 dir, uint32_t depth, const Blob& wvl, Blob& rad, float& pdfS)
{
	if (l.type == PRB_LIGHT_SKY) { // SkyLight::eval, sky.cpp:53-80
		const V3 ld	  = m3mul(l.inv_normal_matrix, dir);
		const float x = (ld.x == 0 && ld.y == 0) ? 1e-5f : ld.x; // Spherical::from_direction
		float az	  = cr_atan2(ld.y, x);
		az			  = az < 0 ? az + 2 * PR_PI : az;
		const float el = 0.5f * PR_PI - cr_acos(ld.z);
		if (!l.sky_extend && el < 0) {
			rad	 = blob(0);
			pdfS = 0;
			return;
		}
		rad	 = skyRadiance(S, l, wvl, el, az);
		pdfS = l.sky_extend ? dist2DContinuousPdf(S, l, fdiv(az, SKY_AZIMUTH_RANGE), fdiv(el, 2 * SKY_ELEVATION_RANGE) + 0.5f)
							: dist2DContinuousPdf(S, l, fdiv(az, SKY_AZIMUTH_RANGE), fdiv(el, SKY_ELEVATION_RANGE));
		const float f	  = cr_cos(el);
		const float denom = 2 * PR_PI * PR_PI * f;
		pdfS *= (denom <= PR_EPSILON) ? 0.0f : 1.0f / denom;
	} else if (l.type == PRB_LIGHT_SUN) { // SunLight::eval, sun.cpp:60-75
		const float cosine = fmaxf(0.0f, dot(dir, ld3(l.sun_dir)));
		if (cosine < l.sun_cos_theta) {
			rad	 = blob(0);
			pdfS = 0;
		} else {
#pragma unroll
			for (int i = 0; i < 4; ++i)
				rad[i] = tableLookup(S.pool + l.table_offset, l.table_count, l.table_start, l.table_end, wvl[i]);
			pdfS = l.sun_pdf;
		}
	} else {
		envEval(S, l, dir, depth, wvl, rad, pdfS);
	}
}
} // namespace prb
