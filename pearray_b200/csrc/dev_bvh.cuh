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
// stack entry .x encodings (.y = entry distance bits, used to cull popped subtrees behind the closest hit)
//   00nn nnnn ...   internal node index
//   1ccf ffff ...   leaf: (count-1) in bits 29..30, first primitive in bits 0..28 (triangle range in a BLAS, entity ref in the TLAS)
//   0xFFFFFFFF      leave the current BLAS (restore the world-space ray)
constexpr uint32_t STK_LEAF	 = 0x80000000u;
constexpr uint32_t STK_EXIT	 = 0xFFFFFFFFu;
constexpr uint32_t STK_NONE	 = 0xFFFFFFFEu;

// float(q) for a byte q without an int->float conversion: 0x4B000000 | q is 2^23 + q
PRB_DEV float byteToFloat(uint32_t word, uint32_t sel) { return __uint_as_float(__byte_perm(word, 0x4B000000u, sel)) - 8388608.0f; }

// Traverses TLAS + BLAS.  ANY: returns at the first accepted primitive.
//
// One loop iteration = at most one internal-node step followed by at most one leaf step, so the lanes of a warp
// re-converge at both phases ("if-if" traversal).  A node step tests the 8 quantised child boxes branch-free, keeps
// the nearest hit child in registers as the next entry and pushes the others; a leaf step runs the watertight test on
// <= 4 triangles (BLAS) or enters an entity (TLAS: analytic sphere, or ray transformed into the mesh's local space).
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
	V3 inv			= mk(safeInv(D.x), safeInv(D.y), safeInv(D.z));
	uint32_t curEnt = PRB_INVALID_ID; // entity whose BLAS is being traversed (TLAS level when invalid)
	uint32_t cur	= S.tlasRoot;
	for (;;) {
		// ---------------------------------------------------------------- internal node step
		if (!(cur & STK_LEAF)) {
			const uint4* np = S.bvhNodes + 5 * (size_t)cur;
			const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
			const float px = __uint_as_float(n0.x), py = __uint_as_float(n0.y), pz = __uint_as_float(n0.z);
			const float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23),
						sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23);
			const uint32_t childBase = n1.x, primBase = n1.y;
			const float tcur = best.t; // == tmax until something was hit
			uint32_t nearEntry = STK_NONE;
			float nearDist	   = PRB_INF;
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
				const bool hit = (meta != 0xFFu) && (fmaxf(tn, tmin) <= fminf(tf, tcur));
				if (hit) {
					const uint32_t entry = (meta & 0x80u) ? (childBase + (meta & 0x7Fu)) : (STK_LEAF | (((meta >> 5) & 3u) << 29) | (primBase + (meta & 0x1Fu)));
					// keep the nearest child in registers, push the other one
					const bool nearer	 = tn < nearDist;
					const uint32_t pe	 = nearer ? nearEntry : entry;
					const float pd		 = nearer ? nearDist : tn;
					nearEntry			 = nearer ? entry : nearEntry;
					nearDist			 = nearer ? tn : nearDist;
					if (pe != STK_NONE && sp < BVH_STACK)
						stack[sp++] = make_uint2(pe, __float_as_uint(pd));
				}
			}
			cur = nearEntry;
		}
		// ---------------------------------------------------------------- leaf step
		if (cur != STK_NONE && (cur & STK_LEAF)) {
			const uint32_t first = cur & 0x1FFFFFFFu, count = ((cur >> 29) & 3u) + 1;
			if (curEnt == PRB_INVALID_ID) {
				// TLAS leaf = one entity (the TLAS is built with one reference per leaf)
				const uint32_t e	 = __ldg(S.tlasRefs + first);
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
					cur = STK_NONE;
				} else {
					// enter the entity's BLAS; the matching exit marker restores the world-space ray
					if (sp < BVH_STACK)
						stack[sp++] = make_uint2(STK_EXIT, 0);
					curEnt = e;
					if (type == PRB_ENTITY_MESH) { // planes are stored in world space: no transform (plane.cpp:71-94)
						O	= xfPoint(en.world_to_local, wO);
						D	= xfVec(en.world_to_local, wD);
						inv = mk(safeInv(D.x), safeInv(D.y), safeInv(D.z));
					}
					cur = en.blas_root;
					continue;
				}
			} else {
				for (uint32_t k = first; k < first + count; ++k) {
					const float4* tp = S.bvhTris + 3 * (size_t)k;
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
				cur = STK_NONE;
			}
		}
		// ---------------------------------------------------------------- pop
		if (cur == STK_NONE) {
			for (;;) {
				if (sp == 0)
					return best.entity != PRB_INVALID_ID;
				const uint2 e = stack[--sp];
				if (e.x == STK_EXIT) { // leave the BLAS: restore the world-space ray
					O	   = wO;
					D	   = wD;
					inv	   = mk(safeInv(D.x), safeInv(D.y), safeInv(D.z));
					curEnt = PRB_INVALID_ID;
					continue;
				}
				if (__uint_as_float(e.y) > best.t)
					continue; // subtree entirely behind the current closest hit
				cur = e.x;
				break;
			}
		}
	}
}
} // namespace prb
