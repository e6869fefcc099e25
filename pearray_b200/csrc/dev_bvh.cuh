// Device scene view and BVH8 traversal (closest hit / any hit).
// Replaces rtcIntersect16 / rtcIntersect1 / rtcOccluded1 (reference src/core/scene/Scene.cpp:138-280).
//
// Layout in HBM: prb_bvh8_node (80 B, five 16-byte loads per node visit), prb_bvh_tri (48 B, three 16-byte
// loads per triangle), two levels: TLAS over entities -> per-mesh BLAS entered with the ray transformed by the
// inverse instance matrix (direction not re-normalised, so t is preserved).
// One thread = one ray; the traversal stack lives in local memory (L1-resident), entries carry the child
// entry distance so popped subtrees behind the current hit are culled.
//
// Closest-hit semantics are ORDER INDEPENDENT: among all primitives accepted by the watertight test inside
// [tmin, tmax] the lexicographically smallest (t, entity, prim) wins, so any correct traversal order gives the
// same (entity, prim, u, v, t) bit for bit.
#pragma once
#include "../../include/prb200_abi.h"
#include "dev_math.cuh"

namespace prb {
struct DScene { // device pointers + by-value small structs; passed to kernels by value
	prb_settings settings;
	prb_camera camera;
	prb_sampler aa, lens, time;
	prb_spectral_mapper mapper;
	const prb_node* nodes;
	const prb_material* materials;
	const prb_emission* emissions;
	const prb_entity* entities;
	const uint32_t* entityMaterials;
	const prb_mesh* meshes;
	const float* vertices;
	const float* normals;
	const float* uvs;
	const uint32_t* faceIndices;
	const uint32_t* faceSlots;
	const prb_light* lights;
	const float* lightCDF;
	const uint4* bvhNodes; // 5 x uint4 per node
	const float4* bvhTris; // 3 x float4 per triangle
	const uint32_t* tlasRefs;
	const float* pool;
	const float* rrProb; // RussianRoulette::probability(len) table
	uint32_t nMaterials, nEmissions, nEntities, nLights, nMeshes, tlasRoot, cieOffset, rrCount;
	uint32_t hasInfLight;
	uint32_t upsamplerOffset, upsamplerRes; // RGB -> spectrum coefficient cube in the pool (image textures)
	uint32_t hasCombined; // any blend / add material in the scene (keeps the check off the path of scenes without them)
	// tiny scenes (<= SMALL_MAX_TRIS triangles in <= SMALL_MAX_ENTS entities: the Cornell boxes and the sphere scene of the
	// reference's examples): a flat list of entity headers + their local-space triangles for the BVH-free trace kernel
	const uint4* small; // nSmallEnts x 4 uint4 headers; the padded world-space boxes of the nSmallFaces faces (lo, hi as float4; lo.w = first triangle, hi.w = triangle count; list padded to a multiple of 8); the triangles (3 x float4 each; c.w = index of the entity header)
	uint32_t nSmallEnts, nSmallFaces, nSmallU4;
	// light path expressions (prb_scene_desc::lpe): dense DFA tables, bytes
	const uint8_t* lpeTables;
	uint32_t nLPE;
	prb_lpe lpe[PRB_MAX_LPE];
};
constexpr uint32_t SMALL_MAX_TRIS = 64, SMALL_MAX_ENTS = 16;

struct HitRec {
	uint32_t entity, prim;
	float u, v, t;
};

PRB_DEV bool betterHit(float t, uint32_t e, uint32_t p, const HitRec& h)
{
	if (h.entity == PRB_INVALID_ID)
		return true;
	if (t != h.t)
		return t < h.t;
	if (e != h.entity)
		return e < h.entity;
	return p < h.prim;
}

// Watertight ray/triangle test (restatement of Embree 3's robust "Pluecker" intersector):
// edge functions relative to the ray origin, accepted when all share a sign within ulp*|U+V+W|, two sided.
// Embree writes cross / dot with msub / madd (common/math/vec3.h), fused multiply-adds in the AVX2 / AVX-512 kernels its ISA
// dispatch selects on current hosts: explicit fmaf in exactly those places, mirrored by std::fma in the oracle.
PRB_DEV float msubE(float a, float b, float c) { return fmaf(a, b, -c); }
PRB_DEV V3 crossE(V3 a, V3 b) { return mk(msubE(a.y, b.z, a.z * b.y), msubE(a.z, b.x, a.x * b.z), msubE(a.x, b.y, a.y * b.x)); }
PRB_DEV float dotE(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
PRB_DEV V3 stableTriangleNormal(V3 a, V3 b, V3 c)
{
	const float ab_x = a.z * b.y, ab_y = a.x * b.z, ab_z = a.y * b.x;
	const float bc_x = b.z * c.y, bc_y = b.x * c.z, bc_z = b.y * c.x;
	const V3 cross_ab = mk(msubE(a.y, b.z, ab_x), msubE(a.z, b.x, ab_y), msubE(a.x, b.y, ab_z));
	const V3 cross_bc = mk(msubE(b.y, c.z, bc_x), msubE(b.z, c.x, bc_y), msubE(b.x, c.y, bc_z));
	const bool sx = fabsf(ab_x) < fabsf(bc_x), sy = fabsf(ab_y) < fabsf(bc_y), sz = fabsf(ab_z) < fabsf(bc_z);
	return mk(sx ? cross_ab.x : cross_bc.x, sy ? cross_ab.y : cross_bc.y, sz ? cross_ab.z : cross_bc.z);
}
PRB_DEV bool triTest(V3 O, V3 D, float tmin, float tmax, V3 p0, V3 p1, V3 p2, float& t, float& u, float& v)
{
	const V3 v0 = p0 - O, v1 = p1 - O, v2 = p2 - O;
	const V3 e0 = v2 - v0, e1 = v0 - v1, e2 = v1 - v2;
	const float U	= dotE(crossE(e0, v2 + v0), D);
	const float V	= dotE(crossE(e1, v0 + v1), D);
	const float W	= dotE(crossE(e2, v1 + v2), D);
	const float UVW = (U + V) + W;
	const float eps = PR_EPSILON * fabsf(UVW);
	const float mn = fminf(U, fminf(V, W)), mx = fmaxf(U, fmaxf(V, W));
	if (!(mn >= -eps || mx <= eps))
		return false;
	const V3 Ng		= stableTriangleNormal(e0, e1, e2);
	const float den = 2 * dotE(Ng, D);
	if (den == 0)
		return false;
	const float T = 2 * dotE(v0, Ng);
	t			  = T / den;
	if (!(tmin <= t && t <= tmax))
		return false;
	if (UVW == 0) {
		u = 0;
		v = 0;
	} else {
		u = fminf(U / UVW, 1.0f);
		v = fminf(V / UVW, 1.0f);
	}
	return true;
}
// analytic sphere (Embree 3 sphere_intersector.h): front hit first, then back hit
PRB_DEV bool sphereTest(V3 O, V3 D, float tmin, float tmax, V3 center, float radius, float& t)
{
	const float rd2	   = 1.0f / dot(D, D);
	const V3 c0		   = center - O;
	const float projC0 = dot(c0, D) * rd2;
	const V3 perp	   = c0 - projC0 * D;
	const float l2	   = dot(perp, perp);
	const float r2	   = radius * radius;
	if (!(l2 <= r2))
		return false;
	const float td		= sqrtf((r2 - l2) * rd2);
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

PRB_DEV V3 xfPoint(const float* m, V3 p)
{
	return mk(((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3], ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7], ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]);
}
PRB_DEV V3 xfVec(const float* m, V3 p)
{
	return mk((m[0] * p.x + m[1] * p.y) + m[2] * p.z, (m[4] * p.x + m[5] * p.y) + m[6] * p.z, (m[8] * p.x + m[9] * p.y) + m[10] * p.z);
}
// the ray into an instance's local space: Embree xfmPoint / xfmVector (common/math/affinespace.h), madd chains
PRB_DEV V3 xfPointE(const float* m, V3 p)
{
	return mk(fmaf(p.x, m[0], fmaf(p.y, m[1], fmaf(p.z, m[2], m[3]))), fmaf(p.x, m[4], fmaf(p.y, m[5], fmaf(p.z, m[6], m[7]))), fmaf(p.x, m[8], fmaf(p.y, m[9], fmaf(p.z, m[10], m[11]))));
}
PRB_DEV V3 xfVecE(const float* m, V3 p)
{
	return mk(fmaf(p.x, m[0], fmaf(p.y, m[1], p.z * m[2])), fmaf(p.x, m[4], fmaf(p.y, m[5], p.z * m[6])), fmaf(p.x, m[8], fmaf(p.y, m[9], p.z * m[10])));
}
PRB_DEV V3 m3mul(const float* m, V3 p)
{
	return mk((m[0] * p.x + m[1] * p.y) + m[2] * p.z, (m[3] * p.x + m[4] * p.y) + m[5] * p.z, (m[6] * p.x + m[7] * p.y) + m[8] * p.z);
}

PRB_DEV float safeInv(float d)
{ // avoid inf * 0 = NaN in the slab test for axis-parallel rays.  The reciprocal only feeds the conservative box test
  // (never a reported t), so the 1-ulp MUFU approximation is enough: its relative error (1.2e-7) scales every plane
  // parameter alike and stays inside the 2e-6 relative slack of the test -- an IEEE division here is ~10 instructions,
  // three times per BLAS entry.
	const float a = fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d);
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
	return r;
}

constexpr int BVH_STACK = 48;
constexpr int BVH_STACK_ALLOC = BVH_STACK + 5; // + the world-space ray (origin, direction, 1/direction, octant) parked while inside a BLAS
// Traversal stack entries are GROUPS (after Ylitie, Karras, Laine: "Efficient Incoherent Ray Traversal on GPUs Through
// Compressed Wide BVHs", HPG 2017), 8 bytes each:
//   node group      .x = child_base of the visited node            .y = hits (8 bit, octant-permuted slot space) | imask << 8
//   primitive group .x = GRP_PRIM | prim_base of the visited node  .y = one bit per primitive of the node's hit leaf children
//                   (BLAS: triangles; TLAS: entity references)
//   exit marker     .x = GRP_EXIT                                  leave the current BLAS, restore the world-space ray
// so a node visit pushes at most ONE entry (the not-yet-visited hit children) instead of up to seven.
constexpr uint32_t GRP_PRIM = 0x80000000u;
constexpr uint32_t GRP_EXIT = 0xFFFFFFFFu;

// float(q) for a byte q without an int->float conversion: 0x4B000000 | q is 2^23 + q
PRB_DEV float byteToFloat(uint32_t word, uint32_t sel) { return __uint_as_float(__byte_perm(word, 0x4B000000u, sel)) - 8388608.0f; }

// ray octant: bit a set when the direction is negative along axis a.  Children are stored by the builder so that
// visiting slots in increasing (slot ^ octant) order is approximately front to back.
PRB_DEV uint32_t rayOctant(V3 inv) { return (inv.x < 0 ? 1u : 0u) | (inv.y < 0 ? 2u : 0u) | (inv.z < 0 ? 4u : 0u); }
// permutes an 8-bit slot mask m so that bit r of the result is bit (r ^ oct) of m
PRB_DEV uint32_t permuteByOctant(uint32_t m, uint32_t oct)
{
	if (oct & 1u)
		m = ((m & 0x55u) << 1) | ((m & 0xAAu) >> 1);
	if (oct & 2u)
		m = ((m & 0x33u) << 2) | ((m & 0xCCu) >> 2);
	if (oct & 4u)
		m = ((m & 0x0Fu) << 4) | ((m & 0xF0u) >> 4);
	return m;
}

// Resumable, WARP-SYNCHRONOUS traversal of TLAS + BLAS: one lane = one ray, but all 32 lanes of the warp call round()
// together (lanes without a ray pass live = false).
//
// round() runs, per lane, a node step (fetch one 80-byte node with five 128-bit loads, test its 8 quantised child boxes
// branch-free, emit a node group + a primitive group), the TLAS primitive phase (analytic sphere, or the ray transformed
// into the mesh's local space and the BLAS entered) and the pop.  Triangle tests are NOT done by the lane that owns the
// ray: the pending (ray, triangle) pairs of the whole warp are counted, and once enough of them are waiting (or no lane has
// node work left) they are dealt out one pair per lane -- the ray travels by shuffle -- so the watertight test runs on full
// warps however few rays reached a leaf in this round (ncu on the 10 M soup: per-lane tests ran at 3.9 of 32 lanes and
// took 47 % of the issue slots).  Accepted hits travel back to the owner by shuffle; closest-hit semantics are order
// independent, so the result is bit-identical.
#ifndef PRB_TRI_BATCH_MIN
#define PRB_TRI_BATCH_MIN 16
#endif
constexpr int TRI_BATCH_MIN = PRB_TRI_BATCH_MIN; // fire a cooperative batch once this many (ray, triangle) pairs are pending in the warp

PRB_DEV uint32_t nthSetBit(uint32_t m, uint32_t r)
{ // position of the r-th (0-based) set bit of m; r < popc(m)
	uint32_t pos = 0, c;
	c = __popc(m & 0xFFFFu);
	if (r >= c) {
		r -= c;
		pos = 16;
		m >>= 16;
	}
	c = __popc(m & 0xFFu);
	if (r >= c) {
		r -= c;
		pos += 8;
		m >>= 8;
	}
	c = __popc(m & 0xFu);
	if (r >= c) {
		r -= c;
		pos += 4;
		m >>= 4;
	}
	c = __popc(m & 0x3u);
	if (r >= c) {
		r -= c;
		pos += 2;
		m >>= 2;
	}
	return pos + ((r >= (m & 1u)) ? 1u : 0u);
}

// The traversal stack (BVH_STACK entries of 8 bytes, local memory) is owned by the caller and passed to round(), so that
// the scalar state below stays in registers.
struct Trav {
	V3 O, D, inv;
	float tmin;
	uint32_t oct, curEnt;
	uint2 ng, pg;
	int sp;
	bool any; // any-hit ray: finished as soon as one primitive was accepted
	HitRec best;

	PRB_DEV void idle()
	{
		ng = pg = make_uint2(0, 0);
		sp		= 0;
		curEnt	= PRB_INVALID_ID;
		any		= false;
		tmin	= 0;
		O = D = inv = mk(0, 0, 0);
		oct					  = 0;
		best.entity			  = PRB_INVALID_ID;
		best.prim			  = 0;
		best.u = best.v = best.t = 0;
	}
	PRB_DEV void begin(const DScene& S, V3 o, V3 d, float t0, float t1, bool anyHit, uint2* __restrict__ stack)
	{

		best.entity = PRB_INVALID_ID;
		best.prim	= 0;
		best.u = best.v = 0;
		best.t			= t1;
		O = o;
		D = d;
		tmin   = t0;
		any	   = anyHit;
		inv	   = mk(safeInv(D.x), safeInv(D.y), safeInv(D.z));
		oct	   = rayOctant(inv);
		curEnt = PRB_INVALID_ID;							   // TLAS level
		// the world-space ray, reloaded (not recomputed: the compiler hoists this path above the pop) when a BLAS is left
		stack[BVH_STACK]	 = make_uint2(__float_as_uint(o.x), __float_as_uint(o.y));
		stack[BVH_STACK + 1] = make_uint2(__float_as_uint(o.z), __float_as_uint(d.x));
		stack[BVH_STACK + 2] = make_uint2(__float_as_uint(d.y), __float_as_uint(d.z));
		stack[BVH_STACK + 3] = make_uint2(__float_as_uint(inv.x), __float_as_uint(inv.y));
		stack[BVH_STACK + 4] = make_uint2(__float_as_uint(inv.z), oct);
		ng	   = make_uint2(S.tlasRoot, (1u << oct) | (1u << 8)); // the root as a one-child node group: slot 0, imask 1
		pg	   = make_uint2(0, 0);
		sp	   = 0;
	}
	PRB_DEV void finish()
	{
		ng.y = 0;
		pg.y = 0;
		sp	 = 0;
	}

	// All 32 lanes must call.  Returns true on the lanes whose ray finished in this round.
	PRB_DEV bool round(const DScene& S, bool live, uint2* __restrict__ stack)
	{
		const unsigned FULL = 0xFFFFFFFFu;
		const int lane		= threadIdx.x & 31;
		bool fin			= false;
		// ---------------------------------------------------------------- node step
		if (live && (ng.y & 0xFFu)) {
			const uint32_t r	= __ffs(ng.y & 0xFFu) - 1; // next child in octant order
			const uint32_t slot = r ^ oct;
			const uint32_t node = ng.x + __popc((ng.y >> 8) & ((1u << slot) - 1u));
			ng.y &= ~(1u << r);
			if ((ng.y & 0xFFu) && sp < BVH_STACK)
				stack[sp++] = ng;
			const uint4* np = S.bvhNodes + 5 * (size_t)node;
			const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
			// Child planes in ray parameter space with ONE fused multiply-add per plane (after Ylitie et al.):
			//   t = (p + q 2^e - O) inv = q adj + org,   adj = 2^e inv (exact),   org = (p - O) inv.
			// The byte q becomes the float F = 2^15 + q with one PRMT (0x47000000 | q << 8), so t = F adj + (org - 2^15 adj).
			// Rounding: |err(t)| <= 4 eps |t| + 2^-9 |adj| (the bias 2^15 adj costs 7 bits relative to one quantisation step);
			// the test stays conservative through the absolute pad 2^-7 |adj| (1/128 step) on the near/far bias and the
			// relative slack 2e-6 in the comparison (rounding above + the ulp slack of the watertight triangle test).
			const float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23),
						sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23);
			const float adjx = sx * inv.x, adjy = sy * inv.y, adjz = sz * inv.z;
			const float bx = fmaf(-32768.0f, adjx, (__uint_as_float(n0.x) - O.x) * inv.x);
			const float by = fmaf(-32768.0f, adjy, (__uint_as_float(n0.y) - O.y) * inv.y);
			const float bz = fmaf(-32768.0f, adjz, (__uint_as_float(n0.z) - O.z) * inv.z);
			const float padx = fabsf(adjx) * 0.0078125f, pady = fabsf(adjy) * 0.0078125f, padz = fabsf(adjz) * 0.0078125f;
			const float bnx = bx - padx, bfx = bx + padx, bny = by - pady, bfy = by + pady, bnz = bz - padz, bfz = bz + padz;
			// near / far plane bytes by the sign of the direction: no per-child min/max
			const bool negx = oct & 1u, negy = oct & 2u, negz = oct & 4u;
			const uint32_t nx0 = negx ? n3.z : n2.x, nx1 = negx ? n3.w : n2.y, fx0 = negx ? n2.x : n3.z, fx1 = negx ? n2.y : n3.w;
			const uint32_t ny0 = negy ? n4.x : n2.z, ny1 = negy ? n4.y : n2.w, fy0 = negy ? n2.z : n4.x, fy1 = negy ? n2.w : n4.y;
			const uint32_t nz0 = negz ? n4.z : n3.x, nz1 = negz ? n4.w : n3.y, fz0 = negz ? n3.x : n4.z, fz1 = negz ? n3.y : n4.w;
			const float tcur  = best.t; // == tmax until something was hit
			uint32_t hitMask  = 0;
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				const uint32_t sel = 0x7404u | ((uint32_t)(i & 3) << 4); // bytes: 00, q, 00, 47  ->  2^15 + q
				const float tnx	   = fmaf(__uint_as_float(__byte_perm(i < 4 ? nx0 : nx1, 0x47000000u, sel)), adjx, bnx);
				const float tny	   = fmaf(__uint_as_float(__byte_perm(i < 4 ? ny0 : ny1, 0x47000000u, sel)), adjy, bny);
				const float tnz	   = fmaf(__uint_as_float(__byte_perm(i < 4 ? nz0 : nz1, 0x47000000u, sel)), adjz, bnz);
				const float tfx	   = fmaf(__uint_as_float(__byte_perm(i < 4 ? fx0 : fx1, 0x47000000u, sel)), adjx, bfx);
				const float tfy	   = fmaf(__uint_as_float(__byte_perm(i < 4 ? fy0 : fy1, 0x47000000u, sel)), adjy, bfy);
				const float tfz	   = fmaf(__uint_as_float(__byte_perm(i < 4 ? fz0 : fz1, 0x47000000u, sel)), adjz, bfz);
				const float tn	   = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
				const float tf	   = fminf(fminf(tfx, tfy), fminf(tfz, tcur));
				hitMask |= (tn <= fmaf(fabsf(tf), 2e-6f, tf)) ? (1u << i) : 0u;
			}
			// children: internal = imask bit, leaf = meta bit 7 clear, empty = meta 0xFF
			const uint32_t imask = n0.w >> 24;
			const uint32_t lz = (~n1.z & 0x80808080u) >> 7, lw = (~n1.w & 0x80808080u) >> 7;
			const uint32_t leafMask = (((lz * 0x01020408u) >> 24) & 0xFu) | (((lw * 0x01020408u) >> 20) & 0xF0u);
			const uint32_t nodeHits = hitMask & imask;
			uint32_t leafHits = hitMask & leafMask, primBits = 0;
			while (leafHits) { // typically 0..2 leaves
				const int i = __ffs(leafHits) - 1;
				leafHits &= leafHits - 1;
				const uint32_t meta = ((i < 4 ? n1.z : n1.w) >> (8 * (i & 3))) & 0xFFu;
				primBits |= ((2u << ((meta >> 5) & 3u)) - 1u) << (meta & 0x1Fu);
			}
			ng = make_uint2(n1.x, permuteByOctant(nodeHits, oct) | ((n0.w >> 24) << 8));
			pg = make_uint2(GRP_PRIM | n1.y, primBits);
		}
		// ---------------------------------------------------------------- TLAS primitive phase (entity references), per lane
		if (live && curEnt == PRB_INVALID_ID) {
			while (pg.y) {
				const uint32_t k = pg.x + (__ffs(pg.y) - 1);
				pg.y &= pg.y - 1;
				const uint32_t e	 = __ldg(S.tlasRefs + (k & ~GRP_PRIM));
				const prb_entity& en = S.entities[e];
				const uint32_t type	 = en.type;
				if (type == PRB_ENTITY_SPHERE) {
					float t;
					if (sphereTest(O, D, tmin, best.t, ld3(en.geo), en.geo[3], t) && betterHit(t, e, 0, best)) {
						best.entity = e;
						best.prim	= 0;
						best.t		= t;
						best.u = best.v = 0;
						if (any) {
							finish();
							fin = true;
						}
					}
				} else {
					// enter the entity's BLAS: park the unfinished TLAS groups under an exit marker
					if ((ng.y & 0xFFu) && sp < BVH_STACK)
						stack[sp++] = ng;
					if (pg.y && sp < BVH_STACK)
						stack[sp++] = pg;
					if (sp < BVH_STACK)
						stack[sp++] = make_uint2(GRP_EXIT, 0);
					curEnt = e;
					if (type == PRB_ENTITY_MESH) { // planes are stored in world space: no transform (plane.cpp:71-94)
						O	= xfPointE(en.world_to_local, O); // at TLAS level (O, D) is the world-space ray
						D	= xfVecE(en.world_to_local, D);
						inv = mk(safeInv(D.x), safeInv(D.y), safeInv(D.z));
						oct = rayOctant(inv);
					}
					ng = make_uint2(en.blas_root, (1u << oct) | (1u << 8));
					pg = make_uint2(0, 0);
				}
			}
		}
		// ---------------------------------------------------------------- cooperative triangle phase (warp-uniform)
		{
			const bool has		 = live && !fin && curEnt != PRB_INVALID_ID && pg.y != 0;
			const bool nodeWork	 = live && !fin && (ng.y & 0xFFu);
			const uint32_t cnt	 = has ? __popc(pg.y) : 0u;
			const unsigned hasM	 = __ballot_sync(FULL, has);
			if (hasM) {
				uint32_t incl = cnt; // inclusive prefix sum of the pair counts
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const uint32_t n = __shfl_up_sync(FULL, incl, d);
					if (lane >= d)
						incl += n;
				}
				const uint32_t total = __shfl_sync(FULL, incl, 31);
				const bool mustNow	 = has && nodeWork && sp >= BVH_STACK; // cannot postpone: no room to park the group
				const bool fire		 = total >= TRI_BATCH_MIN || __ballot_sync(FULL, nodeWork) == 0 || __any_sync(FULL, mustNow);
				if (!fire) {
					if (has && nodeWork) { // park the group, go on with the node group; lanes without node work keep theirs and wait
						stack[sp++] = pg;
						pg.y		= 0;
					}
				} else {
					const uint32_t excl = incl - cnt;
					for (uint32_t base = 0; base < total; base += 32) {
						const uint32_t q = base + lane;
						// owner of pair q: the first lane whose inclusive count exceeds q
						uint32_t o = 0;
#pragma unroll
						for (int step = 16; step >= 1; step >>= 1) {
							const uint32_t v = __shfl_sync(FULL, incl, (o + step - 1) & 31);
							if (v <= q)
								o += step;
						}
						const bool valid = q < total;
						o &= 31;
						const uint32_t r	 = q - __shfl_sync(FULL, excl, o);
						const uint32_t bits	 = __shfl_sync(FULL, pg.y, o);
						const uint32_t pbase = __shfl_sync(FULL, pg.x, o) & ~GRP_PRIM;
						const V3 rO			 = mk(__shfl_sync(FULL, O.x, o), __shfl_sync(FULL, O.y, o), __shfl_sync(FULL, O.z, o));
						const V3 rD			 = mk(__shfl_sync(FULL, D.x, o), __shfl_sync(FULL, D.y, o), __shfl_sync(FULL, D.z, o));
						const float rt0 = __shfl_sync(FULL, tmin, o), rt1 = __shfl_sync(FULL, best.t, o);
						bool hit		= false;
						float t = 0, u = 0, v = 0;
						uint32_t prim = 0;
						if (valid) {
							const uint32_t k = pbase + nthSetBit(bits, r);
							const float4* tp = S.bvhTris + 3 * (size_t)k;
							const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
							hit	 = triTest(rO, rD, rt0, rt1, mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), mk(c.x, c.y, c.z), t, u, v);
							prim = __float_as_uint(a.w);
							if (hit && (__float_as_uint(b.w) & 1u)) {
								u = 1 - u;
								v = 1 - v;
							}
						}
						// accepted hits go back to the ray's lane; the owner keeps the smallest (t, entity, prim)
						unsigned hm = __ballot_sync(FULL, hit);
						while (hm) {
							const int src = __ffs(hm) - 1;
							hm &= hm - 1;
							const uint32_t ow = __shfl_sync(FULL, o, src);
							const float ht	  = __shfl_sync(FULL, t, src);
							const uint32_t hp = __shfl_sync(FULL, prim, src);
							const float hu = __shfl_sync(FULL, u, src), hv = __shfl_sync(FULL, v, src);
							if ((uint32_t)lane == ow && betterHit(ht, curEnt, hp, best)) {
								best.entity = curEnt;
								best.prim	= hp;
								best.t		= ht;
								best.u		= hu;
								best.v		= hv;
							}
						}
					}
					if (has) {
						pg.y = 0;
						if (any && best.entity != PRB_INVALID_ID) {
							finish();
							fin = true;
						}
					}
				}
			}
		}
		// ---------------------------------------------------------------- pop
		if (live && !fin && pg.y == 0) {
			while (!(ng.y & 0xFFu)) {
				if (sp == 0) {
					fin = true;
					break;
				}
				const uint2 e = stack[--sp];
				if (e.x == GRP_EXIT) { // leave the BLAS: restore the world-space ray
					const uint2 w0 = stack[BVH_STACK], w1 = stack[BVH_STACK + 1], w2 = stack[BVH_STACK + 2];
					O	   = mk(__uint_as_float(w0.x), __uint_as_float(w0.y), __uint_as_float(w1.x));
					D	   = mk(__uint_as_float(w1.y), __uint_as_float(w2.x), __uint_as_float(w2.y));
					const uint2 w3 = stack[BVH_STACK + 3], w4 = stack[BVH_STACK + 4];
					inv	   = mk(__uint_as_float(w3.x), __uint_as_float(w3.y), __uint_as_float(w4.x));
					oct	   = w4.y;
					curEnt = PRB_INVALID_ID;
				} else if (e.x & GRP_PRIM) {
					pg = e;
					break; // next round: the node step is skipped (ng is empty), the primitive phases see the group
				} else {
					ng = e;
				}
			}
		}
		return fin;
	}
	PRB_DEV bool hit() const { return best.entity != PRB_INVALID_ID; }
};

// warp-synchronous: every lane of the warp must call (lanes without a ray pass live = false)
PRB_DEV bool traverseScene(const DScene& S, bool live, bool anyHit, V3 wO, V3 wD, float tmin, float tmax, HitRec& best)
{
	Trav tr;
	uint2 stack[BVH_STACK_ALLOC];
	if (live)
		tr.begin(S, wO, wD, tmin, tmax, anyHit, stack);
	else
		tr.idle();
	bool act = live;
	while (__any_sync(0xFFFFFFFFu, act)) {
		if (tr.round(S, act, stack))
			act = false;
	}
	best = tr.best;
	return tr.hit();
}

// BVH-free closest / any hit for tiny scenes (flat list in shared memory, DScene::small), in two phases:
//   1. every lane tests its ray against the padded world-space box of EVERY face (a triangle, or the two triangles of a
//      quad / plane: warp-uniform broadcast loads, ~25 instructions per box, no divergence) and keeps one candidate bit per
//      face whose box meets [tmin, tmax];
//   2. every lane runs the watertight test on the triangles of its own candidates only (1-3 of the 17 faces of a Cornell
//      box), in the triangle's instance-local space like the BVH traversal: same test, same (t, entity, prim) order ->
//      bit-identical results.
// The first version tested every triangle exhaustively (~3400 warp instructions per warp of rays, the second half of each test
// at the few lanes that passed the edge functions; this one ~1700); a BVH traversal of the 32-triangle Cornell box needs ~4900
// at 20 of 32 lanes, almost all of it TLAS -> BLAS entries and exits of eight tiny instances.
// The boxes are conservative in the way the oracle's and the BVH8's are: padded by 8e-6 of the face's largest coordinate
// (covers the rounding of the instance transform and of the plane parameters near the origin), plane parameters widened by
// 4e-6 relative (covers it far away), so that the box test never culls what the triangle test could accept.
// Scenes with at most SMALL_BOX_MIN_FACES faces (the sphere scene: one plane) skip phase 1: every face is a candidate.
constexpr float SMALL_T_SLACK		  = 4e-6f;
constexpr uint32_t SMALL_BOX_MIN_FACES = 4;
// one bit per box (up to 32) that the ray meets within [tmin, tmax]; warp-uniform loop over broadcast loads.  The host pads the
// box list to a multiple of 8 so that the loop unrolls with constant bit positions; the caller masks the padding bits off
PRB_DEV uint32_t smallBoxMask(const float4* __restrict__ boxes, uint32_t n8, V3 inv, V3 oi, float tmin, float tmax)
{
	uint32_t m = 0;
	for (uint32_t g = 0; g < n8; ++g) {
		uint32_t mg = 0;
#pragma unroll
		for (uint32_t k = 0; k < 8; ++k) {
			const float4 lo = boxes[2 * (8 * g + k)], hi = boxes[2 * (8 * g + k) + 1];
			const float ax = fmaf(lo.x, inv.x, oi.x), cx = fmaf(hi.x, inv.x, oi.x);
			const float ay = fmaf(lo.y, inv.y, oi.y), cy = fmaf(hi.y, inv.y, oi.y);
			const float az = fmaf(lo.z, inv.z, oi.z), cz = fmaf(hi.z, inv.z, oi.z);
			float tn	   = fmaxf(fmaxf(fminf(ax, cx), fminf(ay, cy)), fminf(az, cz));
			float tf	   = fminf(fminf(fmaxf(ax, cx), fmaxf(ay, cy)), fmaxf(az, cz));
			tn			   = fmaf(-fabsf(tn), SMALL_T_SLACK, tn); // plane parameters widened (monotonic: once after the reduction is the same as per axis)
			tf			   = fmaf(fabsf(tf), SMALL_T_SLACK, tf);
			if (fmaxf(tn, tmin) <= fminf(tf, tmax))
				mg |= 1u << k;
		}
		m |= mg << (8 * g);
	}
	return m;
}
PRB_DEV uint32_t lowBits(uint32_t n) { return n >= 32u ? 0xFFFFFFFFu : (1u << n) - 1u; }
// warp-synchronous: every lane of the warp must call (lanes without a ray pass live = false); anyHit may differ between lanes
PRB_DEV bool traverseSmall(const uint4* __restrict__ sm, uint32_t nEnts, uint32_t nFaces, bool live, bool anyHit, V3 O, V3 D, float tmin, float tmax, HitRec& best)
{
	best.entity = PRB_INVALID_ID;
	best.prim	= 0;
	best.u = best.v = 0;
	best.t			= tmax;
	const float4* __restrict__ boxes = reinterpret_cast<const float4*>(sm + 4 * nEnts); // per face: lo (w: first triangle), hi (w: triangle count)
	const uint32_t nBoxes			 = (nFaces + 7u) & ~7u;
	const uint4* __restrict__ tris	 = sm + 4 * nEnts + 2 * nBoxes;
	// ---- phase 1: candidate faces by box
	uint32_t cand0 = lowBits(nFaces), cand1 = nFaces > 32u ? lowBits(nFaces - 32u) : 0u;
	if (nFaces > SMALL_BOX_MIN_FACES) { // warp-uniform
		const V3 inv = mk(safeInv(D.x), safeInv(D.y), safeInv(D.z));
		const V3 oi	 = mk(-(O.x * inv.x), -(O.y * inv.y), -(O.z * inv.z));
		cand0 &= smallBoxMask(boxes, (min(nFaces, 32u) + 7u) / 8u, inv, oi, tmin, tmax);
		if (nFaces > 32u)
			cand1 &= smallBoxMask(boxes + 64, (nFaces - 32u + 7u) / 8u, inv, oi, tmin, tmax);
	}
	if (!live)
		cand0 = cand1 = 0u;
	// ---- the analytic spheres (no triangles)
	for (uint32_t e = 0; e < nEnts; ++e) { // warp-uniform
		const uint4 h = sm[4 * e];		   // type, entity id
		if (h.x != PRB_ENTITY_SPHERE)
			continue;
		const float4 r0 = *reinterpret_cast<const float4*>(sm + 4 * e + 1);
		float t;
		if (live && !(anyHit && best.entity != PRB_INVALID_ID) && sphereTest(O, D, tmin, best.t, mk(r0.x, r0.y, r0.z), r0.w, t) && betterHit(t, h.y, 0, best)) {
			best.entity = h.y;
			best.prim	= 0;
			best.t		= t;
			best.u = best.v = 0;
		}
	}
	if (anyHit && best.entity != PRB_INVALID_ID)
		cand0 = cand1 = 0u;
	// ---- phase 2: the watertight test on the triangles of the lane's own candidate faces
	while (__any_sync(0xFFFFFFFFu, (cand0 | cand1) != 0u)) {
		if ((cand0 | cand1) != 0u) {
			uint32_t f;
			if (cand0) {
				f = __ffs(cand0) - 1;
				cand0 &= cand0 - 1;
			} else {
				f = 32 + __ffs(cand1) - 1;
				cand1 &= cand1 - 1;
			}
			const uint32_t first = __float_as_uint(boxes[2 * f].w), count = __float_as_uint(boxes[2 * f + 1].w);
			const float4* tp	 = reinterpret_cast<const float4*>(tris + 3 * first);
			const uint32_t es	 = __float_as_uint(tp[2].w); // index of the face's entity header
			const uint4 h		 = sm[4 * es];
			V3 lo = O, ld = D;
			if (h.x == PRB_ENTITY_MESH) { // planes are stored in world space
				const float4 r0 = *reinterpret_cast<const float4*>(sm + 4 * es + 1), r1 = *reinterpret_cast<const float4*>(sm + 4 * es + 2),
							 r2 = *reinterpret_cast<const float4*>(sm + 4 * es + 3);
				const float m[12] = { r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w };
				lo				  = xfPointE(m, O);
				ld				  = xfVecE(m, D);
			}
#pragma unroll 1
			for (uint32_t j = 0; j < count; ++j, tp += 3) {
				const float4 a = tp[0], b = tp[1], c = tp[2];
				float t, u, v;
				if (triTest(lo, ld, tmin, best.t, mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), mk(c.x, c.y, c.z), t, u, v)) {
					const uint32_t prim = __float_as_uint(a.w);
					if (betterHit(t, h.y, prim, best)) {
						if (__float_as_uint(b.w) & 1u) {
							u = 1 - u;
							v = 1 - v;
						}
						best.entity = h.y;
						best.prim	= prim;
						best.t		= t;
						best.u		= u;
						best.v		= v;
						if (anyHit) {
							cand0 = cand1 = 0u;
							break;
						}
					}
				}
			}
		}
	}
	return best.entity != PRB_INVALID_ID;
}

// warp-aggregated fetch of the next work item from a global counter; every lane of the warp must call it.
// Lanes with `need` get a unique index (>= limit when the pool is exhausted); others get 0xFFFFFFFF.
PRB_DEV uint32_t fetchWork(uint32_t* counter, bool need)
{
	const unsigned mask = __ballot_sync(0xFFFFFFFFu, need);
	if (mask == 0)
		return 0xFFFFFFFFu;
	const int lane	 = threadIdx.x & 31;
	const int leader = __ffs(mask) - 1;
	uint32_t base	 = 0;
	if (lane == leader)
		base = atomicAdd(counter, (uint32_t)__popc(mask));
	base = __shfl_sync(0xFFFFFFFFu, base, leader);
	return need ? base + __popc(mask & ((1u << lane) - 1u)) : 0xFFFFFFFFu;
}
#ifndef PRB_REFILL_LANES
#define PRB_REFILL_LANES 20
#endif
#ifndef PRB_REFILL_LANES_STREAM
#define PRB_REFILL_LANES_STREAM 26
#endif
// A persistent warp leaves the traversal loop to refill its idle lanes when fewer lanes than this are still tracing.
// Measured (gpurun_out/refill_sweep.log): the wavefront k_trace on complex.prc is flat between 12 and 26 (20 best by 1 %);
// the ray-stream kernels on the 10 M-triangle soup gain steadily with the threshold (12: 933, 20: 966, 26: 1007 Mrays/s primary).
constexpr int REFILL_LANES		  = PRB_REFILL_LANES;
constexpr int REFILL_LANES_STREAM = PRB_REFILL_LANES_STREAM;
} // namespace prb
