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

constexpr int BVH_STACK = 96;
constexpr uint32_t STACK_INSTANCE_EXIT = 0xFFFFFFFEu;

// Traverses TLAS + BLAS.  ANY: returns at the first accepted primitive.  `nodeVisits`/`triTests` are optional
// work counters (profiling builds).
template <bool ANY>
PRB_DEV bool traverseScene(const DScene& S, V3 wO, V3 wD, float tmin, float tmax, HitRec& best)
{
	best.entity = PRB_INVALID_ID;
	best.prim	= 0;
	best.u = best.v = 0;
	best.t			= tmax;
	uint2 stack[BVH_STACK];
	int sp = 0;
	V3 O = wO, D = wD;
	V3 inv			 = mk(safeInv(D.x), safeInv(D.y), safeInv(D.z));
	uint32_t curEnt	 = PRB_INVALID_ID; // entity whose BLAS is being traversed
	bool found		 = false;
	uint32_t node	 = S.tlasRoot;
	for (;;) {
		// ---- visit internal node `node`
		const uint4* np = S.bvhNodes + 5 * (size_t)node;
		const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
		const float px = __uint_as_float(n0.x), py = __uint_as_float(n0.y), pz = __uint_as_float(n0.z);
		const float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23),
					sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23);
		const uint32_t childBase = n1.x, primBase = n1.y;
		const uint32_t metaLo = n1.z, metaHi = n1.w;
		const float tcur = found ? best.t : tmax;
		// near-first ordering: collect internal hits, leaves are intersected immediately
		uint32_t hitNode[8];
		float hitDist[8];
		int nh = 0;
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			const uint32_t meta = ((i < 4 ? metaLo : metaHi) >> (8 * (i & 3))) & 0xFFu;
			if (meta == 0xFFu)
				continue;
			const uint32_t sh = 8 * (i & 3);
			const uint32_t qlx = ((i < 4 ? n2.x : n2.y) >> sh) & 0xFFu, qly = ((i < 4 ? n2.z : n2.w) >> sh) & 0xFFu;
			const uint32_t qlz = ((i < 4 ? n3.x : n3.y) >> sh) & 0xFFu, qhx = ((i < 4 ? n3.z : n3.w) >> sh) & 0xFFu;
			const uint32_t qhy = ((i < 4 ? n4.x : n4.y) >> sh) & 0xFFu, qhz = ((i < 4 ? n4.z : n4.w) >> sh) & 0xFFu;
			const float lox = px + (float)qlx * sx, loy = py + (float)qly * sy, loz = pz + (float)qlz * sz;
			const float hix = px + (float)qhx * sx, hiy = py + (float)qhy * sy, hiz = pz + (float)qhz * sz;
			const float ax = (lox - O.x) * inv.x, bx = (hix - O.x) * inv.x;
			const float ay = (loy - O.y) * inv.y, by = (hiy - O.y) * inv.y;
			const float az = (loz - O.z) * inv.z, bz = (hiz - O.z) * inv.z;
			float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
			float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
			tn		 = tn - fabsf(tn) * 4e-7f; // conservative: never cull what the triangle test could accept
			tf		 = tf + fabsf(tf) * 4e-7f;
			if (!(fmaxf(tn, tmin) <= fminf(tf, tcur)))
				continue;
			if (meta & 0x80u) {
				hitNode[nh] = childBase + (meta & 0x7Fu);
				hitDist[nh] = tn;
				++nh;
			} else {
				const uint32_t first = primBase + (meta & 0x1Fu), count = ((meta >> 5) & 3u) + 1;
				for (uint32_t k = first; k < first + count; ++k) {
					if (curEnt == PRB_INVALID_ID) {
						// ---- TLAS leaf: an entity
						const uint32_t e	 = __ldg(S.tlasRefs + k);
						const prb_entity& en = S.entities[e];
						if (en.type == PRB_ENTITY_SPHERE) {
							float t;
							if (sphereTest(O, D, tmin, found ? best.t : tmax, ld3(en.geo), en.geo[3], t)) {
								if (betterHit(t, e, 0, best)) {
									best.entity = e;
									best.prim	= 0;
									best.t		= t;
									best.u = best.v = 0;
								}
								found = true;
								if (ANY)
									return true;
							}
						} else if (sp + 2 <= BVH_STACK) {
							// defer: push the instance (entered when popped); distance = this leaf box entry
							stack[sp++] = make_uint2(0x80000000u | e, __float_as_uint(tn));
						}
					} else {
						const float4* tp = S.bvhTris + 3 * (size_t)k;
						const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
						float t, u, v;
						if (triTest(O, D, tmin, found ? best.t : tmax, mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), mk(c.x, c.y, c.z), t, u, v)) {
							const uint32_t prim = __float_as_uint(a.w);
							if (__float_as_uint(b.w) & 1u) {
								u = 1 - u;
								v = 1 - v;
							}
							if (betterHit(t, curEnt, prim, best)) {
								best.entity = curEnt;
								best.prim	= prim;
								best.t		= t;
								best.u		= u;
								best.v		= v;
							}
							found = true;
							if (ANY)
								return true;
						}
					}
				}
			}
		}
		// push internal hits far-to-near (selection by insertion sort on <= 8 entries)
		for (int i = 1; i < nh; ++i) {
			const float d	 = hitDist[i];
			const uint32_t n = hitNode[i];
			int j			 = i - 1;
			while (j >= 0 && hitDist[j] < d) {
				hitDist[j + 1] = hitDist[j];
				hitNode[j + 1] = hitNode[j];
				--j;
			}
			hitDist[j + 1] = d;
			hitNode[j + 1] = n;
		}
		for (int i = 0; i < nh && sp < BVH_STACK; ++i)
			stack[sp++] = make_uint2(hitNode[i], __float_as_uint(hitDist[i]));
		// ---- pop
		bool haveNode = false;
		while (sp > 0) {
			const uint2 e = stack[--sp];
			if (e.x == STACK_INSTANCE_EXIT) { // leave the BLAS: restore the world-space ray
				O	   = wO;
				D	   = wD;
				inv	   = mk(safeInv(D.x), safeInv(D.y), safeInv(D.z));
				curEnt = PRB_INVALID_ID;
				continue;
			}
			if (found && __uint_as_float(e.y) > best.t)
				continue; // subtree entirely behind the current closest hit
			if (e.x & 0x80000000u) { // enter instance
				const uint32_t ent	 = e.x & 0x7FFFFFFFu;
				const prb_entity& en = S.entities[ent];
				stack[sp++]			 = make_uint2(STACK_INSTANCE_EXIT, 0);
				curEnt				 = ent;
				if (en.type == PRB_ENTITY_MESH) {
					O = xfPoint(en.world_to_local, wO);
					D = xfVec(en.world_to_local, wD);
				} // planes are stored in world space: no transform (plane.cpp:71-94)
				inv		 = mk(safeInv(D.x), safeInv(D.y), safeInv(D.z));
				node	 = en.blas_root;
				haveNode = true;
				break;
			}
			node	 = e.x;
			haveNode = true;
			break;
		}
		if (!haveNode)
			break;
	}
	return found;
}
} // namespace prb
