// Host SAH BVH builder producing the compressed 8-wide layout of include/prb200_abi.h (prb_bvh8_node).
// Replaces Embree's rtcCommitScene (reference src/core/scene/Scene.cpp:99-101, mesh.cpp:113-118).
//   1. binned-SAH binary BVH (16 bins, parallel over large subtrees, leaves <= maxLeafPrims),
//   2. greedy collapse to 8 children (repeatedly open the child with the largest surface area),
//   3. child boxes quantised to 8 bit on a per-node power-of-two grid, rounded OUTWARDS and verified against the
//      exact decode the device uses, so the compressed boxes are conservative.
#include "prh.h"

#include <atomic>
#include <future>
#include <numeric>
#include <thread>

namespace PR {
namespace {
struct Node2 {
	BoundingBox box;
	int32 left = -1, right = -1; // children (internal) ...
	uint32 first = 0, count = 0; // ... or prim range (leaf) in the index array
	bool leaf() const { return left < 0; }
};

struct Builder2 {
	const std::vector<BoundingBox>& boxes;
	std::vector<Vector3f> centers;
	std::vector<uint32> indices;
	int maxLeaf;
	std::vector<Node2> nodes; // arena; subtree builders append under a mutex-free scheme: pre-reserved 2N
	std::atomic<uint32> nodeCounter{ 0 };
	std::atomic<int> tasksInFlight{ 0 };
	int maxTasks;

	Builder2(const std::vector<BoundingBox>& b, int ml)
		: boxes(b)
		, maxLeaf(ml)
	{
		const size_t n = boxes.size();
		centers.resize(n);
		indices.resize(n);
		for (size_t i = 0; i < n; ++i) {
			centers[i] = boxes[i].center();
			indices[i] = (uint32)i;
		}
		nodes.resize(std::max<size_t>(1, 2 * n));
		maxTasks = (int)std::max(1u, std::thread::hardware_concurrency());
	}
	uint32 alloc() { return nodeCounter.fetch_add(1); }

	void build(uint32 nodeIdx, uint32 first, uint32 count)
	{
		BoundingBox box, cbox;
		for (uint32 i = first; i < first + count; ++i) {
			box.combine(boxes[indices[i]]);
			cbox.combine(centers[indices[i]]);
		}
		Node2& node = nodes[nodeIdx];
		node.box	= box;
		if ((int)count <= maxLeaf) {
			node.first = first;
			node.count = count;
			return;
		}
		// binned SAH over the largest centroid axes
		constexpr int BINS = 16;
		int bestAxis	   = -1, bestSplit = -1;
		float bestCost	   = PR_INF;
		const Vector3f ext = cbox.hi - cbox.lo;
		for (int axis = 0; axis < 3; ++axis) {
			if (!(ext[axis] > 0))
				continue;
			BoundingBox bb[BINS];
			uint32 bc[BINS]	  = {};
			const float scale = BINS / ext[axis];
			for (uint32 i = first; i < first + count; ++i) {
				const uint32 p = indices[i];
				int b		   = (int)((centers[p][axis] - cbox.lo[axis]) * scale);
				b			   = std::min(BINS - 1, std::max(0, b));
				bb[b].combine(boxes[p]);
				bc[b]++;
			}
			float rightArea[BINS];
			uint32 rightCount[BINS];
			BoundingBox acc;
			uint32 cnt = 0;
			for (int b = BINS - 1; b > 0; --b) {
				if (bc[b])
					acc.combine(bb[b]);
				cnt += bc[b];
				rightArea[b]  = cnt ? acc.halfArea() : 0;
				rightCount[b] = cnt;
			}
			acc = BoundingBox();
			cnt = 0;
			for (int b = 0; b < BINS - 1; ++b) {
				if (bc[b])
					acc.combine(bb[b]);
				cnt += bc[b];
				if (cnt == 0 || rightCount[b + 1] == 0)
					continue;
				const float cost = acc.halfArea() * cnt + rightArea[b + 1] * rightCount[b + 1];
				if (cost < bestCost) {
					bestCost  = cost;
					bestAxis  = axis;
					bestSplit = b;
				}
			}
		}
		uint32 mid;
		if (bestAxis < 0) { // all centroids coincide: split the index range in the middle
			mid = first + count / 2;
		} else {
			const float scale = BINS / ext[bestAxis];
			auto* begin		  = indices.data() + first;
			auto* it		  = std::partition(begin, begin + count, [&](uint32 p) {
				 int b = (int)((centers[p][bestAxis] - cbox.lo[bestAxis]) * scale);
				 b	   = std::min(BINS - 1, std::max(0, b));
				 return b <= bestSplit;
			 });
			mid				  = first + (uint32)(it - begin);
			if (mid == first || mid == first + count)
				mid = first + count / 2;
		}
		const uint32 l = alloc(), r = alloc();
		nodes[nodeIdx].left	 = (int32)l;
		nodes[nodeIdx].right = (int32)r;
		const uint32 lc = mid - first, rc = first + count - mid;
		if (count > 65536 && tasksInFlight.load() < maxTasks) {
			tasksInFlight++;
			auto fut = std::async(std::launch::async, [this, l, first, lc] { build(l, first, lc); });
			build(r, mid, rc);
			fut.get();
			tasksInFlight--;
		} else {
			build(l, first, lc);
			build(r, mid, rc);
		}
	}
};

inline float decodeScale(uint8 e)
{
	const uint32 bits = (uint32)e << 23;
	float f;
	std::memcpy(&f, &bits, 4);
	return f;
}
// device decode: p + q * scale, fp32, separate multiply and add (no FMA contraction)
inline float decodeCoord(float p, uint8 q, float scale)
{
	volatile float prod = (float)q * scale;
	volatile float sum	= p + prod;
	return sum;
}
} // namespace

BVH8 buildBVH8(const BVHBuildInput& in, int maxLeafPrims)
{
	BVH8 out;
	const size_t n = in.boxes.size();
	maxLeafPrims   = std::max(1, std::min(4, maxLeafPrims));
	if (n == 0) {
		prb_bvh8_node root{};
		std::memset(root.meta, 0xFF, 8);
		root.ex = root.ey = root.ez = 1;
		out.nodes.push_back(root);
		return out;
	}
	Builder2 b2(in.boxes, maxLeafPrims);
	const uint32 root2 = b2.alloc();
	b2.build(root2, 0, (uint32)n);
	out.bounds = b2.nodes[root2].box;

	// ---- collapse: BFS so that the internal children of a node are contiguous
	struct Work {
		uint32 node8;
		int32 node2;
	};
	out.nodes.reserve(n / 4 + 8);
	out.primOrder.reserve(n);
	out.nodes.emplace_back();
	std::vector<Work> queue;
	queue.push_back({ 0, (int32)root2 });
	size_t head = 0;
	while (head < queue.size()) {
		const Work w = queue[head++];
		// gather up to 8 children
		int32 kids[8];
		int nk			= 0;
		const Node2& me = b2.nodes[w.node2];
		if (me.leaf()) { // root is a leaf (tiny input): single leaf child
			kids[nk++] = w.node2;
		} else {
			kids[nk++] = me.left;
			kids[nk++] = me.right;
			while (nk < 8) {
				int best	   = -1;
				float bestArea = -1;
				for (int i = 0; i < nk; ++i) {
					const Node2& c = b2.nodes[kids[i]];
					if (c.leaf())
						continue;
					const float a = c.box.halfArea();
					if (a > bestArea) {
						bestArea = a;
						best	 = i;
					}
				}
				if (best < 0)
					break;
				const Node2& c = b2.nodes[kids[best]];
				kids[best]	   = c.left;
				kids[nk++]	   = c.right;
			}
		}
		prb_bvh8_node node{};
		std::memset(node.meta, 0xFF, 8);
		BoundingBox nb;
		for (int i = 0; i < nk; ++i)
			nb.combine(b2.nodes[kids[i]].box);
		// Slot assignment for octant-ordered traversal (Ylitie et al. 2017, section 3.2): the child placed in slot s should
		// lie towards the diagonal d_s = (s&1 ? + : -, s&2 ? + : -, s&4 ? + : -) of the node, so that a ray with octant o
		// (bit a set = negative direction on axis a) meets the slots in increasing (s ^ o) order roughly front to back.
		// Greedy maximisation of sum_c dot(centroid_c - centroid_node, d_slot(c)).
		{
			const Vector3f nc = nb.center();
			float score[8][8];
			for (int c = 0; c < nk; ++c) {
				const Vector3f d = b2.nodes[kids[c]].box.center() - nc;
				for (int sl = 0; sl < 8; ++sl)
					score[c][sl] = ((sl & 1) ? d.x : -d.x) + ((sl & 2) ? d.y : -d.y) + ((sl & 4) ? d.z : -d.z);
			}
			int32 slotKid[8];
			bool kidDone[8] = {}, slotDone[8] = {};
			for (int sl = 0; sl < 8; ++sl)
				slotKid[sl] = -1;
			for (int round = 0; round < nk; ++round) {
				int bc = -1, bs = -1;
				float bv = -PR_INF;
				for (int c = 0; c < nk; ++c) {
					if (kidDone[c])
						continue;
					for (int sl = 0; sl < 8; ++sl)
						if (!slotDone[sl] && score[c][sl] > bv) {
							bv = score[c][sl];
							bc = c;
							bs = sl;
						}
				}
				kidDone[bc]	 = true;
				slotDone[bs] = true;
				slotKid[bs]	 = kids[bc];
			}
			for (int sl = 0; sl < 8; ++sl)
				kids[sl] = slotKid[sl];
		}
		node.px = nb.lo.x;
		node.py = nb.lo.y;
		node.pz = nb.lo.z;
		uint8 ebits[3];
		float scale[3];
		for (int a = 0; a < 3; ++a) {
			const float extent = nb.hi[a] - nb.lo[a];
			int e			   = -126;
			if (extent > 0) {
				e = (int)std::ceil(std::log2((double)extent / 255.0));
				while (std::ldexp(255.0, e) < (double)extent)
					++e;
				e = std::max(-126, std::min(127, e));
			}
			ebits[a] = (uint8)(e + 127);
			scale[a] = decodeScale(ebits[a]);
		}
		node.ex			= ebits[0];
		node.ey			= ebits[1];
		node.ez			= ebits[2];
		node.prim_base = (uint32)out.primOrder.size();
		// internal children are appended to out.nodes while this node is processed, so they are contiguous
		uint32 firstChildIndex = 0;
		uint32 internalCount   = 0;
		uint32 primOffset	   = 0;
		uint8 imask			   = 0;
		for (int i = 0; i < 8; ++i) {
			if (kids[i] < 0)
				continue; // empty slot: meta stays 0xFF
			const Node2& c = b2.nodes[kids[i]];
			const float clo[3] = { c.box.lo.x, c.box.lo.y, c.box.lo.z }, chi[3] = { c.box.hi.x, c.box.hi.y, c.box.hi.z };
			const float p[3] = { node.px, node.py, node.pz };
			uint8 qlo[3], qhi[3];
			for (int a = 0; a < 3; ++a) {
				int lo = (int)std::floor(((double)clo[a] - (double)p[a]) / (double)scale[a]);
				int hi = (int)std::ceil(((double)chi[a] - (double)p[a]) / (double)scale[a]);
				lo	   = std::max(0, std::min(255, lo));
				hi	   = std::max(0, std::min(255, hi));
				while (lo > 0 && decodeCoord(p[a], (uint8)lo, scale[a]) > clo[a])
					--lo;
				while (hi < 255 && decodeCoord(p[a], (uint8)hi, scale[a]) < chi[a])
					++hi;
				qlo[a] = (uint8)lo;
				qhi[a] = (uint8)hi;
			}
			node.qlo_x[i] = qlo[0];
			node.qlo_y[i] = qlo[1];
			node.qlo_z[i] = qlo[2];
			node.qhi_x[i] = qhi[0];
			node.qhi_y[i] = qhi[1];
			node.qhi_z[i] = qhi[2];
			if (c.leaf()) {
				node.meta[i] = (uint8)(((c.count - 1) << 5) | primOffset);
				for (uint32 k = 0; k < c.count; ++k)
					out.primOrder.push_back(b2.indices[c.first + k]);
				primOffset += c.count;
			} else {
				const uint32 idx = (uint32)out.nodes.size();
				if (internalCount == 0)
					firstChildIndex = idx;
				out.nodes.emplace_back();
				queue.push_back({ idx, kids[i] });
				node.meta[i] = (uint8)(0x80 | internalCount);
				imask |= (uint8)(1u << i);
				++internalCount;
			}
		}
		node.child_base	   = firstChildIndex;
		node.imask		   = imask;
		out.nodes[w.node8] = node;
	}
	return out;
}
} // namespace PR
