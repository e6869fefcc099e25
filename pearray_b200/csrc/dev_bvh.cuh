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
	uint32_t hasEnvLight;
};

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
PRB_DEV V3 stableTriangleNormal(V3 a, V3 b, V3 c)
{
	const float ab_x = a.z * b.y, ab_y = a.x * b.z, ab_z = a.y * b.x;
	const float bc_x = b.z * c.y, bc_y = b.x * c.z, bc_z = b.y * c.x;
	const V3 cross_ab = mk(a.y * b.z - ab_x, a.z * b.x - ab_y, a.x * b.y - ab_z);
	const V3 cross_bc = mk(b.y * c.z - bc_x, b.z * c.x - bc_y, b.x * c.y - bc_z);
	const bool sx = fabsf(ab_x) < fabsf(bc_x), sy = fabsf(ab_y) < fabsf(bc_y), sz = fabsf(ab_z) < fabsf(bc_z);
	return mk(sx ? cross_ab.x : cross_bc.x, sy ? cross_ab.y : cross_bc.y, sz ? cross_ab.z : cross_bc.z);
}
PRB_DEV bool triTest(V3 O, V3 D, float tmin, float tmax, V3 p0, V3 p1, V3 p2, float& t, float& u, float& v)
{
	const V3 v0 = p0 - O, v1 = p1 - O, v2 = p2 - O;
	const V3 e0 = v2 - v0, e1 = v0 - v1, e2 = v1 - v2;
	const float U	= dot(cross(e0, v2 + v0), D);
	const float V	= dot(cross(e1, v0 + v1), D);
	const float W	= dot(cross(e2, v1 + v2), D);
	const float UVW = (U + V) + W;
	const float eps = PR_EPSILON * fabsf(UVW);
	const float mn = fminf(U, fminf(V, W)), mx = fmaxf(U, fmaxf(V, W));
	if (!(mn >= -eps || mx <= eps))
		return false;
	const V3 Ng		= stableTriangleNormal(e0, e1, e2);
	const float den = 2 * dot(Ng, D);
	if (den == 0)
		return false;
	const float T = 2 * dot(v0, Ng);
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
PRB_DEV V3 m3mul(const float* m, V3 p)
{
	return mk((m[0] * p.x + m[1] * p.y) + m[2] * p.z, (m[3] * p.x + m[4] * p.y) + m[5] * p.z, (m[6] * p.x + m[7] * p.y) + m[8] * p.z);
}

PRB_DEV float safeInv(float d)
{ // avoid inf * 0 = NaN in the slab test for axis-parallel rays
	const float a = fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d);
	return 1.0f / a;
}

constexpr int BVH_STACK = 48;
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

// Resumable traversal of TLAS + BLAS for ONE ray (one thread = one ray).
//
// advance() runs one round: a node step (fetch one 80-byte node with five 128-bit loads, test its 8 quantised child
// boxes branch-free, emit a node group + a primitive group), the primitive phase (watertight triangle tests in a BLAS;
// entity entry -- analytic sphere or ray transformed into the mesh's local space -- in the TLAS) and the pop.  Keeping the
// state in a struct lets the persistent kernels leave the loop to refill idle lanes with new rays and come back.
// Primitive groups are POSTPONED (pushed) while only a few lanes of the warp have primitives to test and node work is
// available, so triangle tests run with fuller warps.
constexpr int TRI_POSTPONE_LANES = 10;

struct Trav {
	V3 wO, wD, O, D, inv;
	float tmin;
	uint32_t oct, curEnt;
	uint2 ng, pg;
	int sp;
	HitRec best;
	uint2 stack[BVH_STACK];

	PRB_DEV void begin(const DScene& S, V3 o, V3 d, float t0, float t1)
	{
		best.entity = PRB_INVALID_ID;
		best.prim	= 0;
		best.u = best.v = 0;
		best.t			= t1;
		wO = O = o;
		wD = D = d;
		tmin   = t0;
		inv	   = mk(safeInv(D.x), safeInv(D.y), safeInv(D.z));
		oct	   = rayOctant(inv);
		curEnt = PRB_INVALID_ID;							   // TLAS level
		ng	   = make_uint2(S.tlasRoot, (1u << oct) | (1u << 8)); // the root as a one-child node group: slot 0, imask 1
		pg	   = make_uint2(0, 0);
		sp	   = 0;
	}

	// returns true when the ray is finished (ANY: as soon as one primitive was accepted)
	template <bool ANY>
	PRB_DEV bool advance(const DScene& S)
	{
		// ---------------------------------------------------------------- node step
		if (ng.y & 0xFFu) {
			const uint32_t r	= __ffs(ng.y & 0xFFu) - 1; // next child in octant order
			const uint32_t slot = r ^ oct;
			const uint32_t node = ng.x + __popc((ng.y >> 8) & ((1u << slot) - 1u));
			ng.y &= ~(1u << r);
			if ((ng.y & 0xFFu) && sp < BVH_STACK)
				stack[sp++] = ng;
			const uint4* np = S.bvhNodes + 5 * (size_t)node;
			const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
			const float px = __uint_as_float(n0.x), py = __uint_as_float(n0.y), pz = __uint_as_float(n0.z);
			const float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23),
						sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23);
			const float tcur  = best.t; // == tmax until something was hit
			uint32_t nodeHits = 0, primBits = 0;
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				const uint32_t meta = ((i < 4 ? n1.z : n1.w) >> (8 * (i & 3))) & 0xFFu;
				const uint32_t sel	= 0x7440u + (uint32_t)(i & 3); // byte (i&3) of the first operand into the low mantissa byte
				// child box: lo = p + q_lo * 2^e (the product is exact, so the fused form rounds like the builder's decodeCoord)
				const float lox = fmaf(byteToFloat(i < 4 ? n2.x : n2.y, sel), sx, px), loy = fmaf(byteToFloat(i < 4 ? n2.z : n2.w, sel), sy, py);
				const float loz = fmaf(byteToFloat(i < 4 ? n3.x : n3.y, sel), sz, pz), hix = fmaf(byteToFloat(i < 4 ? n3.z : n3.w, sel), sx, px);
				const float hiy = fmaf(byteToFloat(i < 4 ? n4.x : n4.y, sel), sy, py), hiz = fmaf(byteToFloat(i < 4 ? n4.z : n4.w, sel), sz, pz);
				const float ax = (lox - O.x) * inv.x, bx = (hix - O.x) * inv.x;
				const float ay = (loy - O.y) * inv.y, by = (hiy - O.y) * inv.y;
				const float az = (loz - O.z) * inv.z, bz = (hiz - O.z) * inv.z;
				float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
				float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
				tn		 = tn - fabsf(tn) * 4e-7f; // conservative: never cull what the triangle test could accept
				tf		 = tf + fabsf(tf) * 4e-7f;
				const bool hit			= (meta != 0xFFu) && (fmaxf(tn, tmin) <= fminf(tf, tcur));
				const uint32_t leafBits = ((2u << ((meta >> 5) & 3u)) - 1u) << (meta & 0x1Fu);
				nodeHits |= (hit && (meta & 0x80u)) ? (1u << i) : 0u;
				primBits |= (hit && !(meta & 0x80u)) ? leafBits : 0u;
			}
			ng = make_uint2(n1.x, permuteByOctant(nodeHits, oct) | ((n0.w >> 24) << 8));
			pg = make_uint2(GRP_PRIM | n1.y, primBits);
		}
		// ---------------------------------------------------------------- primitive phase
		while (pg.y) {
			if ((ng.y & 0xFFu) && sp < BVH_STACK && __popc(__activemask()) < TRI_POSTPONE_LANES) {
				stack[sp++] = pg; // too few lanes have primitives: test them later, go on with the node group
				pg.y		= 0;
				break;
			}
			const uint32_t k = pg.x + (__ffs(pg.y) - 1); // GRP_PRIM | primitive index
			pg.y &= pg.y - 1;
			if (curEnt == PRB_INVALID_ID) {
				// TLAS: one entity per reference
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
						if (ANY)
							return true;
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
						O	= xfPoint(en.world_to_local, wO);
						D	= xfVec(en.world_to_local, wD);
						inv = mk(safeInv(D.x), safeInv(D.y), safeInv(D.z));
						oct = rayOctant(inv);
					}
					ng = make_uint2(en.blas_root, (1u << oct) | (1u << 8));
					pg = make_uint2(0, 0);
				}
			} else {
				const float4* tp = S.bvhTris + 3 * (size_t)(k & ~GRP_PRIM);
				const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
				float t, u, v;
				if (triTest(O, D, tmin, best.t, mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), mk(c.x, c.y, c.z), t, u, v)) {
					const uint32_t prim = __float_as_uint(a.w);
					if (betterHit(t, curEnt, prim, best)) {
						if (__float_as_uint(b.w) & 1u) {
							u = 1 - u;
							v = 1 - v;
						}
						best.entity = curEnt;
						best.prim	= prim;
						best.t		= t;
						best.u		= u;
						best.v		= v;
						if (ANY)
							return true;
					}
				}
			}
		}
		// ---------------------------------------------------------------- pop
		while (!(ng.y & 0xFFu)) {
			if (sp == 0)
				return true;
			const uint2 e = stack[--sp];
			if (e.x == GRP_EXIT) { // leave the BLAS: restore the world-space ray
				O	   = wO;
				D	   = wD;
				inv	   = mk(safeInv(D.x), safeInv(D.y), safeInv(D.z));
				oct	   = rayOctant(inv);
				curEnt = PRB_INVALID_ID;
			} else if (e.x & GRP_PRIM) {
				pg = e;
				break; // next round: the node step is skipped (ng is empty), the primitive phase runs
			} else {
				ng = e;
			}
		}
		return false;
	}
	PRB_DEV bool hit() const { return best.entity != PRB_INVALID_ID; }
};

template <bool ANY>
PRB_DEV bool traverseScene(const DScene& S, V3 wO, V3 wD, float tmin, float tmax, HitRec& best)
{
	Trav tr;
	tr.begin(S, wO, wD, tmin, tmax);
	while (!tr.advance<ANY>(S)) {
	}
	best = tr.best;
	return tr.hit();
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
constexpr int REFILL_LANES = 20; // leave the traversal loop to refill idle lanes when fewer lanes than this are still tracing
} // namespace prb
