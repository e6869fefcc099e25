// Host-side simulator of the device BVH8 traversal (pearray_b200/csrc/dev_bvh.cuh) for design decisions without a GPU:
// counts node visits, wasted visits (no child hit), triangle tests and stack traffic per ray on the C5 soup for traversal
// variants (group entry-distance culling, nearest-child-first).  Not bit-exact with the device; statistics only.
//   g++ -O2 -std=c++17 tools/bvh_sim.cpp -Iinclude -Lpearray_b200 -lprb200_host -lprb200 -Wl,-rpath,$PWD/pearray_b200 -o /tmp/bvh_sim
#include "prb200_abi.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
extern "C" {
void* prh_make_soup(uint32_t, uint64_t, uint32_t, uint32_t);
const prb_scene_desc* prh_scene_desc(void*);
}
struct V3 { float x, y, z; };
static V3 sub(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
static V3 add(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
static V3 mul(V3 a, float f) { return { a.x * f, a.y * f, a.z * f }; }
static float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
static bool tri(V3 O, V3 D, float tmin, float tmax, const prb_bvh_tri& T, float& t)
{
	const V3 p0{ T.v0[0], T.v0[1], T.v0[2] }, p1{ T.v1[0], T.v1[1], T.v1[2] }, p2{ T.v2[0], T.v2[1], T.v2[2] };
	const V3 v0 = sub(p0, O), v1 = sub(p1, O), v2 = sub(p2, O);
	const V3 e0 = sub(v2, v0), e1 = sub(v0, v1), e2 = sub(v1, v2);
	const float U = dot(cross(e0, add(v2, v0)), D), V = dot(cross(e1, add(v0, v1)), D), W = dot(cross(e2, add(v1, v2)), D);
	const float UVW = U + V + W, eps = 1.1920929e-7f * std::fabs(UVW);
	const float mn = std::min(U, std::min(V, W)), mx = std::max(U, std::max(V, W));
	if (!(mn >= -eps || mx <= eps)) return false;
	const V3 Ng = cross(e2, e1);
	const float den = 2 * dot(Ng, D);
	if (den == 0) return false;
	t = 2 * dot(v0, Ng) / den;
	return tmin <= t && t <= tmax;
}
struct Stats { double nodes = 0, wasted = 0, tris = 0, pushes = 0, culledGroups = 0, culledChildren = 0, rays = 0, hits = 0; };
enum Variant { BASE = 0, GROUP_REST = 1, NEAREST_FIRST = 2, PER_CHILD = 3, SORTED = 4 };
// returns closest t (or tmax)
static float trace(const prb_scene_desc* d, V3 O, V3 D, float tmin, float tmax, int variant, Stats& st)
{
	struct Entry { uint32_t base; uint32_t hits, imask; float bound; float tn[8]; };
	std::vector<Entry> stack;
	const V3 inv{ 1.0f / (std::fabs(D.x) > 1e-20f ? D.x : 1e-20f), 1.0f / (std::fabs(D.y) > 1e-20f ? D.y : 1e-20f), 1.0f / (std::fabs(D.z) > 1e-20f ? D.z : 1e-20f) };
	const uint32_t oct = (inv.x < 0 ? 1 : 0) | (inv.y < 0 ? 2 : 0) | (inv.z < 0 ? 4 : 0);
	float best = tmax;
	const uint32_t root = d->entities[0].blas_root;
	Entry cur{ root, 1u, 1u, 0, {} }; // hits in SLOT space here (bit i = child i)
	cur.tn[0] = 0;
	bool rootGroup = true;
	for (;;) {
		while (cur.hits == 0) {
			if (stack.empty()) return best;
			cur = stack.back();
			stack.pop_back();
			rootGroup = false;
			if (variant == GROUP_REST || variant == NEAREST_FIRST) {
				if (cur.bound > best * (1 + 2e-6f)) { st.culledGroups++; st.culledChildren += __builtin_popcount(cur.hits); cur.hits = 0; }
			} else if (variant == PER_CHILD || variant == SORTED) {
				for (int i = 0; i < 8; ++i)
					if ((cur.hits >> i & 1) && cur.tn[i] > best * (1 + 2e-6f)) { cur.hits &= ~(1u << i); st.culledChildren++; }
			}
		}
		// choose next child: octant order = increasing (slot ^ oct)
		int pick = -1;
		if (variant == SORTED) {
			float b = INFINITY;
			for (int i = 0; i < 8; ++i) if ((cur.hits >> i & 1) && cur.tn[i] < b) { b = cur.tn[i]; pick = i; }
		} else {
			for (int r = 0; r < 8 && pick < 0; ++r) if (cur.hits >> (r ^ oct) & 1) pick = r ^ oct;
		}
		const uint32_t node = rootGroup ? cur.base : cur.base + __builtin_popcount(cur.imask & ((1u << pick) - 1u));
		cur.hits &= ~(1u << pick);
		if (cur.hits) { stack.push_back(cur); st.pushes++; }
		rootGroup = false;
		// node step
		const prb_bvh8_node& n = d->bvh_nodes[node];
		st.nodes++;
		const float sx = std::ldexp(1.0f, n.ex - 127), sy = std::ldexp(1.0f, n.ey - 127), sz = std::ldexp(1.0f, n.ez - 127);
		Entry next{ n.child_base, 0, n.imask, INFINITY, {} };
		uint32_t leafHits = 0;
		float tnAll[8];
		for (int i = 0; i < 8; ++i) {
			tnAll[i] = INFINITY;
			if (n.meta[i] == 0xFF) continue;
			const float lox = n.px + n.qlo_x[i] * sx, hix = n.px + n.qhi_x[i] * sx, loy = n.py + n.qlo_y[i] * sy, hiy = n.py + n.qhi_y[i] * sy,
						loz = n.pz + n.qlo_z[i] * sz, hiz = n.pz + n.qhi_z[i] * sz;
			const float tx0 = (lox - O.x) * inv.x, tx1 = (hix - O.x) * inv.x, ty0 = (loy - O.y) * inv.y, ty1 = (hiy - O.y) * inv.y, tz0 = (loz - O.z) * inv.z,
						tz1 = (hiz - O.z) * inv.z;
			const float tn = std::max(std::max(std::min(tx0, tx1), std::min(ty0, ty1)), std::max(std::min(tz0, tz1), tmin));
			const float tf = std::min(std::min(std::max(tx0, tx1), std::max(ty0, ty1)), std::min(std::max(tz0, tz1), best));
			if (tn <= tf * (1 + 2e-6f)) {
				tnAll[i] = tn;
				if (n.meta[i] & 0x80) next.hits |= 1u << i; else leafHits |= 1u << i;
			}
		}
		if (!next.hits && !leafHits) st.wasted++;
		// leaves: test now (the device batches them; the count is what matters)
		for (int i = 0; i < 8; ++i)
			if (leafHits >> i & 1) {
				const uint32_t cnt = ((n.meta[i] >> 5) & 3) + 1, first = n.prim_base + (n.meta[i] & 0x1F);
				for (uint32_t k = 0; k < cnt; ++k) {
					float t;
					st.tris++;
					if (tri(O, D, tmin, best, d->bvh_tris[first + k], t)) best = t;
				}
			}
		// next group
		for (int i = 0; i < 8; ++i) next.tn[i] = tnAll[i];
		if (variant == GROUP_REST && next.hits) { // bound over all but the first child in octant order
			int first = -1;
			for (int r = 0; r < 8 && first < 0; ++r) if (next.hits >> (r ^ oct) & 1) first = r ^ oct;
			float b = INFINITY;
			for (int i = 0; i < 8; ++i) if ((next.hits >> i & 1) && i != first) b = std::min(b, tnAll[i]);
			next.bound = b;
		}
		if (variant == NEAREST_FIRST && next.hits) { // visit the nearest child right away, bound = second smallest
			int a = -1; float m1 = INFINITY, m2 = INFINITY;
			for (int i = 0; i < 8; ++i) if (next.hits >> i & 1) { if (tnAll[i] < m1) { m2 = m1; m1 = tnAll[i]; a = i; } else m2 = std::min(m2, tnAll[i]); }
			next.bound = m2;
			// emulate: visit a first -> handled by giving it priority: push remainder, continue with single-child group
			Entry rest = next; rest.hits &= ~(1u << a);
			if (rest.hits) { stack.push_back(rest); st.pushes++; }
			next.hits = 1u << a;
		}
		// truncate bound to bf16 like the device
		if (std::isfinite(next.bound)) { uint32_t u; std::memcpy(&u, &next.bound, 4); u &= 0xFFFF0000u; std::memcpy(&next.bound, &u, 4); }
		cur = next;
	}
}
int main(int argc, char** argv)
{
	const uint32_t ntri = argc > 1 ? atoi(argv[1]) : 10000000;
	const int nrays = argc > 2 ? atoi(argv[2]) : 20000;
	void* h = prh_make_soup(ntri, 1234, 2048, 2048);
	const prb_scene_desc* d = prh_scene_desc(h);
	std::mt19937 rng(7);
	std::uniform_real_distribution<float> U(0, 1);
	std::vector<V3> O1, D1, O2, D2;
	const float half = std::tan(20.0f * 3.14159265f / 180);
	for (int i = 0; i < nrays; ++i) {
		const float x = (2 * U(rng) - 1) * half, y = (2 * U(rng) - 1) * half;
		const float l = std::sqrt(x * x + y * y + 1);
		O1.push_back({ 0, 0, -3 });
		D1.push_back({ x / l, y / l, 1 / l });
	}
	const char* names[] = { "base (octant order, no cull)", "group bound over rest", "nearest first + bound", "per-child cull at pop (octant order)", "per-child cull + sorted order" };
	for (int pass = 0; pass < 2; ++pass) {
		const auto& Os = pass ? O2 : O1; const auto& Ds = pass ? D2 : D1;
		printf("== %s rays (%zu)\n", pass ? "incoherent bounce" : "primary", Os.size());
		for (int v = 0; v < 5; ++v) {
			Stats st;
			for (size_t i = 0; i < Os.size(); ++i) {
				const float t = trace(d, Os[i], Ds[i], pass ? 1e-4f : 1e-6f, INFINITY, v, st);
				st.rays++;
				if (std::isfinite(t)) {
					st.hits++;
					if (!pass && v == 0) {
						V3 P = add(Os[i], mul(Ds[i], t));
						V3 w{ 2 * U(rng) - 1, 2 * U(rng) - 1, 2 * U(rng) - 1 };
						const float l = std::sqrt(dot(w, w));
						O2.push_back(P); D2.push_back(mul(w, 1 / l));
					}
				}
			}
			printf("%-40s nodes/ray %6.1f wasted %5.1f%% tris/ray %6.1f pushes/ray %5.1f culled groups/ray %5.1f children %5.1f hit %.3f\n", names[v], st.nodes / st.rays,
				   100 * st.wasted / st.nodes, st.tris / st.rays, st.pushes / st.rays, st.culledGroups / st.rays, st.culledChildren / st.rays, st.hits / st.rays);
		}
	}
	return 0;
}
