// Wavefront kernels of the 'direct' path tracer (reference src/plugins/main/integrators/direct.cpp:44-473,
// src/vcm/vcm/Walker.h:24-55) restructured for the GPU:
//
//   slot == one film pixel owned by this context.  A pixel's samples are generated strictly one after the other
//   because all decisions of a pixel draw from the pixel's own pcg32_fast stream (RenderRandomMap), carried
//   across iterations.  Slots are independent, so the wave always holds (#owned pixels) paths in flight:
//   a slot whose path ended is refilled IN PLACE with its next sample ("path regeneration"), which keeps every
//   lane busy until the pixel ran out of samples and makes slot == thread: all state accesses are coalesced SoA
//   float4 loads/stores, no queue indirection.
//
//   one wavefront iteration = 2 kernels
//     k_trace  per slot: the pending NEE shadow ray (any hit; adds its contribution, and finishes the previous sample
//              when that sample ended with the shadow ray still in flight), then the path's next ray (closest hit)
//     k_shade  per slot: emission / environment on the hit or miss, NEE (light sample + material eval + MIS -> shadow
//              ray), Russian roulette + material sample -> next ray; when the path ends: fold the sample into the film
//              (running mean) and start the pixel's next camera sample
//   Slots retire when their pixel has no samples left; a device counter tells the host when all have retired.
#pragma once
#include "dev_shade.cuh"

namespace prb {
enum { CNT_RETIRED = 0, CNT_WORK = 1, CNT__COUNT = 4 };
enum { ST_CAMERA_RAY = 0, ST_LIGHT_RAY, ST_PRIMARY, ST_BOUNCE, ST_SHADOW, ST_MONO, ST_PIXEL_SAMPLE, ST_ENTITY_HIT, ST_BG_HIT, ST_CAMERA_DEPTH, ST_LIGHT_DEPTH, ST__COUNT };

constexpr uint32_t FD_DEPTH_MASK   = 0xFFFFu;
constexpr uint32_t FD_FLAGS_SHIFT  = 16; // ray flags (8 bit)
constexpr uint32_t FD_LAST_DELTA   = 1u << 30;
constexpr uint32_t FD_LAST_EMISSIVE = 1u << 31;

// slot state bits
constexpr uint32_t SF_ACTIVE   = 1u; // rayO/rayD hold a ray to extend
constexpr uint32_t SF_SHADOW   = 2u; // shO/shD/shXYZ hold a pending NEE shadow ray
constexpr uint32_t SF_FINALIZE = 4u; // the sample that spawned the shadow ray already ended: fold prevAcc (+ contribution) into the film

struct WFState {
	// per slot
	uint32_t* pixel;
	uint32_t* iter;	 // number of samples started so far (== index of the next iteration to generate)
	uint32_t* state; // SF_* bits
	float4* rayO;	 // xyz, tmin
	float4* rayD;	 // xyz, tmax
	float4* wvl;
	uint32_t* flagsDepth;
	float4* thr;
	float4* pathPDF;
	float4* prevPDF;
	float4* wvlPDF;
	float4* lastPos;
	uint4* hit; // entity, prim, u bits, v bits
	float* hitT;
	float4* shO;	 // shadow origin xyz, tmin
	float4* shD;	 // shadow dir xyz, tmax
	float4* shXYZ;	 // contribution if visible
	float4* iterXYZ; // XYZ accumulated by the running sample
	float4* prevAcc; // XYZ of the sample waiting for its last shadow ray, w = its 1-based iteration count
	uint32_t* counters;
	// film (indexed by film pixel)
	uint64_t* rng;
	float* filmMean; // 3 per pixel, running mean, unfiltered
	uint32_t* sampleCount;
	float* aov; // 10 per pixel or null
	unsigned long long* stats;
	uint32_t nSlots, firstIter, endIter;
};

PRB_DEV void statAdd(unsigned long long* stats, int which, uint32_t v)
{ // one atomic per warp per counter (all lanes call)
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
	if ((threadIdx.x & 31) == 0 && v)
		atomicAdd(stats + which, (unsigned long long)v);
}

// FrameOutputDevice::onEndOfIteration (FrameOutputDevice.cpp:202-221): film = (film * (i - 1) + sample) / i
PRB_DEV void foldSampleIntoFilm(const WFState& W, uint32_t pix, float x, float y, float z, uint32_t iterCount)
{
	const float fin	  = (float)(iterCount - 1);
	const float iterf = (float)iterCount;
	float* m		  = W.filmMean + 3 * (size_t)pix;
	m[0]			  = (m[0] * fin + x) / iterf;
	m[1]			  = (m[1] * fin + y) / iterf;
	m[2]			  = (m[2] * fin + z) / iterf;
}

// starts the pixel's next camera sample in `slot` (RenderTile::constructCameraRay); returns false when the pixel has
// no samples left (the slot retires)
PRB_DEV bool startNextSample(const DScene& S, const WFState& W, uint32_t slot, uint32_t pix)
{
	const uint32_t it = W.iter[slot];
	if (it >= W.endIter)
		return false;
	Rng rnd{ W.rng[pix] };
	CameraSampleOut cs;
	const uint32_t fw = S.settings.film_width;
	constructCameraRay(S, pix % fw, pix / fw, it, rnd, cs);
	W.rng[pix]		   = rnd.s;
	W.iter[slot]	   = it + 1;
	W.rayO[slot]	   = make_float4(cs.origin.x, cs.origin.y, cs.origin.z, cs.tmin);
	W.rayD[slot]	   = make_float4(cs.dir.x, cs.dir.y, cs.dir.z, cs.tmax);
	W.wvl[slot]		   = tof4(cs.wvl);
	W.wvlPDF[slot]	   = tof4(cs.wvlPDF);
	W.thr[slot]		   = make_float4(1, 1, 1, 1);
	W.pathPDF[slot]	   = make_float4(1, 1, 1, 1);
	W.prevPDF[slot]	   = make_float4(1, 1, 1, 1);
	W.lastPos[slot]	   = make_float4(0, 0, 0, 0);
	const uint32_t rf  = PRB_RAY_CAMERA | (cs.mono ? PRB_RAY_MONOCHROME : 0);
	W.flagsDepth[slot] = (rf << FD_FLAGS_SHIFT) | FD_LAST_DELTA; // depth 0, LastWasDelta = true
	return true;
}

__global__ void __launch_bounds__(128) k_init_slots(const __grid_constant__ DScene S, WFState W)
{
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t started	= 0;
	if (slot < W.nSlots) {
		W.iter[slot]	= W.firstIter;
		W.iterXYZ[slot] = make_float4(0, 0, 0, 0);
		const bool ok	= startNextSample(S, W, slot, W.pixel[slot]);
		W.state[slot]	= ok ? SF_ACTIVE : 0u;
		started			= ok ? 1u : 0u;
		if (!ok)
			atomicAdd(W.counters + CNT_RETIRED, 1u);
	}
	if (blockIdx.x == 0 && threadIdx.x == 0)
		W.counters[CNT_WORK] = 0;
	statAdd(W.stats, ST_PIXEL_SAMPLE, started);
	statAdd(W.stats, ST_CAMERA_RAY, started);
	statAdd(W.stats, ST_PRIMARY, started);
}

// ------------------------------------------------------------------ trace: pending shadow ray, then the path's next ray
// Persistent threads: lanes pull slots from a global counter (warp-aggregated atomic); a lane whose ray finished leaves
// the traversal loop and, once fewer than REFILL_LANES lanes of the warp are still tracing, the warp refills its idle lanes
// with new slots, so divergence in traversal length does not idle lanes until the longest ray of a warp is done.
// static variant (slot == thread): no work counter, no refill.  Cheaper for scenes whose rays finish within a few steps
// (a handful of primitives), where the bookkeeping of the persistent variant costs more than the idle lanes it avoids.
__global__ void __launch_bounds__(128) k_trace_static(DScene S, WFState W)
{
	// The traversal is warp-synchronous (lanes without a ray take part with live = false), so idle lanes cost issue
	// slots.  Each phase therefore first compacts the block's slots that have a ray of that kind (not every active slot
	// has a shadow ray; slots whose pixel ran out of samples have none) into a list in shared memory and traces the list
	// with dense warps; warps past the end of the list leave immediately.  (Measured on the Cornell box: neutral -- idle
	// slots there already come in whole warps because neighbouring pixels retire together -- kept for scenes where they do not.)
	__shared__ uint16_t list[128];
	__shared__ uint32_t count[2];
	const uint32_t base = blockIdx.x * blockDim.x;
	const uint32_t own	= base + threadIdx.x;
	const uint32_t st0	= own < W.nSlots ? W.state[own] : 0u;
	if (threadIdx.x < 2)
		count[threadIdx.x] = 0;
	__syncthreads();
	if (st0 & SF_SHADOW)
		list[atomicAdd(&count[0], 1u)] = (uint16_t)threadIdx.x;
	__syncthreads();
	{
		const bool live		= threadIdx.x < count[0];
		const uint32_t slot = base + (live ? list[threadIdx.x] : 0u);
		const uint32_t st	= live ? W.state[slot] : 0u;
		float4 o = make_float4(0, 0, 0, 0), d = make_float4(0, 0, 1, 0);
		if (live) {
			o = W.shO[slot];
			d = W.shD[slot];
		}
		HitRec h;
		const bool occluded = traverseScene(S, live, true, mk(o.x, o.y, o.z), mk(d.x, d.y, d.z), o.w, d.w, h);
		if (live) {
			const float4 c = occluded ? make_float4(0, 0, 0, 0) : W.shXYZ[slot];
			if (st & SF_FINALIZE) {
				const float4 p = W.prevAcc[slot];
				foldSampleIntoFilm(W, W.pixel[slot], p.x + c.x, p.y + c.y, p.z + c.z, __float_as_uint(p.w));
			} else if (!occluded) {
				float4 acc = W.iterXYZ[slot];
				acc.x += c.x;
				acc.y += c.y;
				acc.z += c.z;
				W.iterXYZ[slot] = acc;
			}
			W.state[slot] = st & SF_ACTIVE;
		}
	}
	__syncthreads(); // the list is rebuilt for the path rays (the SF_ACTIVE bit of a slot is not touched by the shadow phase)
	if (st0 & SF_ACTIVE)
		list[atomicAdd(&count[1], 1u)] = (uint16_t)threadIdx.x;
	__syncthreads();
	{
		const bool live		= threadIdx.x < count[1];
		const uint32_t slot = base + (live ? list[threadIdx.x] : 0u);
		float4 o = make_float4(0, 0, 0, 0), d = make_float4(0, 0, 1, 0);
		if (live) {
			o = W.rayO[slot];
			d = W.rayD[slot];
		}
		HitRec h;
		traverseScene(S, live, false, mk(o.x, o.y, o.z), mk(d.x, d.y, d.z), o.w, d.w, h);
		if (live) {
			W.hit[slot]	 = make_uint4(h.entity, h.prim, __float_as_uint(h.u), __float_as_uint(h.v));
			W.hitT[slot] = h.t;
		}
	}
}

__global__ void __launch_bounds__(128) k_trace(DScene S, WFState W)
{
	Trav tr;
	uint2 stack[BVH_STACK_ALLOC];
	tr.idle();
	uint32_t slot = 0, st = 0;
	int phase	  = 0; // 0 idle, 1 shadow ray in flight, 2 closest-hit ray in flight
	bool pool	  = true;
	for (;;) {
		// ---- refill idle lanes (warp-uniform loop: a fetched slot may turn out to be retired, then the lane asks again)
		for (;;) {
			const bool need = (phase == 0) && pool;
			if (!__any_sync(0xFFFFFFFFu, need))
				break;
			const uint32_t s = fetchWork(W.counters + CNT_WORK, need);
			if (need) {
				if (s >= W.nSlots) {
					pool = false;
				} else {
					slot = s;
					st	 = W.state[slot];
					if (st & SF_SHADOW) {
						const float4 o = W.shO[slot], d = W.shD[slot];
						tr.begin(S, mk(o.x, o.y, o.z), mk(d.x, d.y, d.z), o.w, d.w, true, stack);
						phase = 1;
					} else if (st & SF_ACTIVE) {
						const float4 o = W.rayO[slot], d = W.rayD[slot];
						tr.begin(S, mk(o.x, o.y, o.z), mk(d.x, d.y, d.z), o.w, d.w, false, stack);
						phase = 2;
					}
				}
			}
		}
		if (__ballot_sync(0xFFFFFFFFu, phase != 0) == 0)
			break;
		// ---- trace: warp-synchronous rounds (Trav::round), left to refill once too few lanes are still tracing
		for (;;) {
			const bool done = tr.round(S, phase != 0, stack);
			if (done) {
				if (phase == 1) {
					const bool occluded = tr.hit();
					const float4 c		= occluded ? make_float4(0, 0, 0, 0) : W.shXYZ[slot];
					if (st & SF_FINALIZE) {
						const float4 p = W.prevAcc[slot];
						foldSampleIntoFilm(W, W.pixel[slot], p.x + c.x, p.y + c.y, p.z + c.z, __float_as_uint(p.w));
					} else if (!occluded) {
						float4 acc = W.iterXYZ[slot];
						acc.x += c.x;
						acc.y += c.y;
						acc.z += c.z;
						W.iterXYZ[slot] = acc;
					}
					W.state[slot] = st & SF_ACTIVE;
					if (st & SF_ACTIVE) { // the same lane goes on with the slot's path ray
						const float4 o = W.rayO[slot], d = W.rayD[slot];
						tr.begin(S, mk(o.x, o.y, o.z), mk(d.x, d.y, d.z), o.w, d.w, false, stack);
						phase = 2;
					} else {
						phase = 0;
					}
				} else {
					W.hit[slot]	 = make_uint4(tr.best.entity, tr.best.prim, __float_as_uint(tr.best.u), __float_as_uint(tr.best.v));
					W.hitT[slot] = tr.best.t;
					phase		 = 0;
				}
			}
			const unsigned am = __ballot_sync(0xFFFFFFFFu, phase != 0);
			if (am == 0 || (__popc(am) < REFILL_LANES && __any_sync(0xFFFFFFFFu, pool && phase == 0)))
				break;
		}
	}
}

// ------------------------------------------------------------------ shade
PRB_DEV float misTerm(bool power, float a) { return power ? a * a : a; } // vcm/MIS.h:7-30
PRB_DEV Blob misTermB(bool power, Blob a) { return power ? a * a : a; }

// LocalFrameOutputDevice::commitSpectrals2 (LocalFrameOutputDevice.cpp:88-164) for one fragment; the pixel filter
// is applied as a linear post pass.  Returns false when the fragment is rejected (NaN / Inf / negative).
PRB_DEV bool fragmentXYZ(const DScene& S, const Blob& mis, const Blob& importance, const Blob& radiance, uint32_t rayFlags, bool groupMono, const Blob& grpWvl,
						 float xyz[3])
{
	const bool isMono		  = rayFlags & PRB_RAY_MONOCHROME;
	const Blob heroFactor	  = isMono ? heroOnly() : blob(1);
	const Blob grpImportance  = groupMono ? heroOnly() : blob(1); // CameraRay Importance (* HeroOnly when monochrome), RenderTile.cpp:124-128
	const Blob imp			  = grpImportance * importance;
	const Blob contrib		  = heroFactor * ((mis * imp) * radiance);
	bool invalid			  = false;
#pragma unroll
	for (int i = 0; i < 4; ++i)
		if (isinf(contrib[i]) || isnan(contrib[i]) || contrib[i] < -PR_EPSILON)
			invalid = true;
	if (invalid)
		return false;
	xyz[0] = xyz[1] = xyz[2] = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
#pragma unroll
		for (int c = 0; c < 3; ++c)
			xyz[c] += contrib[k] * cieEval(S, c, grpWvl[k]);
	}
	// BlendWeight == 1 for all supported spectral mappers
	return true;
}

PRB_DEV float rrProbability(const DScene& S, uint32_t pathLength, bool delta)
{ // RussianRoulette::probability, vcm/RussianRoulette.h:22-34 (table computed by the host in double like std::pow)
	if (pathLength == 0 || delta)
		return 1.0f;
	return __ldg(S.rrProb + min(pathLength, S.rrCount - 1));
}

// Slots of one block are dealt to its threads SORTED BY MATERIAL (counting sort in shared memory over the block's window of
// SHADE_BLOCK slots), so that the lanes of a warp run the same material code: k_shade is divergence and instruction-fetch
// bound on scenes with several material types (ncu on boltsandgears: 7.6 of 32 lanes per instruction, 24 warps stalled on
// instruction fetch per issue).  Which thread shades a slot does not change its result: every decision of a pixel draws
// from the pixel's own RNG stream.
// Two instantiations: SHADE_BLOCK_UNIFORM threads and a one-pass window for scenes whose materials all share one type
// (nothing to gain from larger windows; small blocks balance better), SHADE_BLOCK_MIXED threads sorting a window of up to
// SHADE_ROUNDS_MIXED * SHADE_BLOCK_MIXED slots, shaded in that many passes, for scenes that mix material types
// (boltsandgears: 1006 ms -> 447 ms of k_shade per 64 spp).
#ifndef PRB_SHADE_BLOCK_UNIFORM
#define PRB_SHADE_BLOCK_UNIFORM 128
#endif
constexpr int SHADE_BLOCK_UNIFORM = PRB_SHADE_BLOCK_UNIFORM;
constexpr int SHADE_BLOCK_MIXED	  = 512;
constexpr int SHADE_ROUNDS_MIXED	  = 4;
constexpr int SHADE_BINS		  = 64; // materials 0..61 (ids beyond share bin 61), 62 = miss, 63 = no work
PRB_DEV uint32_t shadeSortKey(const DScene& S, const WFState& W, uint32_t slot)
{
	if (slot >= W.nSlots || !(W.state[slot] & SF_ACTIVE))
		return SHADE_BINS - 1;
	const uint4 h = W.hit[slot];
	if (h.x == PRB_INVALID_ID)
		return SHADE_BINS - 2;
	const prb_entity& en = S.entities[h.x];
	uint32_t fslot		 = 0;
	if (en.type == PRB_ENTITY_MESH && en.material_count > 1)
		fslot = S.faceSlots[S.meshes[en.mesh_id].face_offset + h.y];
	const uint32_t mat = fslot < en.material_count ? S.entityMaterials[en.material_offset + fslot] : 0u;
	return min(mat, (uint32_t)SHADE_BINS - 3);
}

#ifndef PRB_SHADE_MINB
#define PRB_SHADE_MINB 4 /* resident 128-thread blocks per SM the uniform instantiation is compiled for */
#endif
template <int SHADE_BLOCK, int SHADE_ROUNDS_MAX, int MATERIALS>
__global__ void __launch_bounds__(SHADE_BLOCK, SHADE_BLOCK <= 128 ? (PRB_SHADE_MINB * 128) / SHADE_BLOCK : 512 / SHADE_BLOCK) k_shade(const __grid_constant__ DScene S, WFState W, int roundsArg)
{
	const int rounds = SHADE_ROUNDS_MAX == 1 ? 1 : roundsArg; // compile-time 1 for the uniform instantiation: no loop

	const prb_settings& st = S.settings;
	const bool power	   = st.mis_power;
	uint32_t sEntity = 0, sBg = 0, sDepth = 0, sShadow = 0, sBounce = 0, sMono = 0;
	uint32_t sSamples = 0;
	__shared__ uint32_t binCount[SHADE_BINS], binStart[SHADE_BINS];
	__shared__ uint16_t order[SHADE_BLOCK * SHADE_ROUNDS_MAX];
	__shared__ uint16_t regenList[SHADE_BLOCK];
	__shared__ uint32_t regenCount[SHADE_ROUNDS_MAX];
	const uint32_t base = blockIdx.x * (uint32_t)(rounds * SHADE_BLOCK);
	{
		if (threadIdx.x < SHADE_BINS)
			binCount[threadIdx.x] = 0;
		if (threadIdx.x < SHADE_ROUNDS_MAX)
			regenCount[threadIdx.x] = 0;
		__syncthreads();
		uint32_t key[SHADE_ROUNDS_MAX], rank[SHADE_ROUNDS_MAX];
#pragma unroll
		for (int r = 0; r < SHADE_ROUNDS_MAX; ++r)
			if (r < rounds) {
				key[r]	= shadeSortKey(S, W, base + r * SHADE_BLOCK + threadIdx.x);
				rank[r] = atomicAdd(&binCount[key[r]], 1u);
			}
		__syncthreads();
		if (threadIdx.x < 32) { // exclusive prefix sum over the 64 bins by one warp (two bins per lane)
			const uint32_t a = binCount[2 * threadIdx.x], b = binCount[2 * threadIdx.x + 1];
			uint32_t incl = a + b;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d);
				if ((int)threadIdx.x >= d)
					incl += n;
			}
			binStart[2 * threadIdx.x]	  = incl - a - b;
			binStart[2 * threadIdx.x + 1] = incl - b;
		}
		__syncthreads();
#pragma unroll
		for (int r = 0; r < SHADE_ROUNDS_MAX; ++r)
			if (r < rounds)
				order[binStart[key[r]] + rank[r]] = (uint16_t)(r * SHADE_BLOCK + threadIdx.x);
		__syncthreads();
	}
#pragma unroll 1
	for (int round = 0; round < rounds; ++round) {
		const uint32_t slot = base + order[round * SHADE_BLOCK + threadIdx.x];
		bool pushShadow		= false;
		bool regen			= false;
		const uint32_t sst	= slot < W.nSlots ? W.state[slot] : 0u;
		if (sst & SF_ACTIVE) {
			const uint32_t pix	= W.pixel[slot];
			const uint4 hraw	= W.hit[slot];
			const float4 ro = W.rayO[slot], rd = W.rayD[slot];
			const uint32_t fd	= W.flagsDepth[slot];
			const uint32_t depth = fd & FD_DEPTH_MASK;
			const uint32_t rayFlags = (fd >> FD_FLAGS_SHIFT) & 0xFFu;
			const Blob wvl		= blob4(W.wvl[slot]);
			const bool groupMono = st.spectral_mono || !st.spectral_hero;
			const V3 O			= mk(ro.x, ro.y, ro.z);
			V3 D				= mk(rd.x, rd.y, rd.z);
			if (depth == 0)
				D = normalized(D); // RayStream::getRay re-normalises on read, RayStream.cpp:167
			float4 accv = W.iterXYZ[slot];
			float acc[3] = { accv.x, accv.y, accv.z };
			Blob Throughput = blob4(W.thr[slot]), PathPDF = blob4(W.pathPDF[slot]), PrevPathPDF = blob4(W.prevPDF[slot]);
			const Blob WavelengthPDF = blob4(W.wvlPDF[slot]);
			bool LastWasDelta = fd & FD_LAST_DELTA, LastWasEmissive = fd & FD_LAST_EMISSIVE;
			bool alive = false;
			float xyz[3];

			if (hraw.x == PRB_INVALID_ID) {
				// ---------------- miss
				++sBg;
				if (depth == 0) { // IntegratorUtils::handleBackgroundGroup, IntegratorUtils.h:16-53
					++sDepth;
					bool illuminated = false;
					for (uint32_t i = 0; i < S.nLights; ++i) {
						const prb_light& l = S.lights[i];
						if (!isInfLight(l) || isDeltaLight(l))
							continue;
						illuminated = true;
						Blob rad;
						float pdfS;
						infLightEval(S, l, D, depth, wvl, rad, pdfS);
						if (fragmentXYZ(S, blob(1), blob(1), rad, rayFlags, groupMono, wvl, xyz)) {
							acc[0] += xyz[0];
							acc[1] += xyz[1];
							acc[2] += xyz[2];
						}
					}
					(void)illuminated; // a zero fragment adds nothing
				} else { // handleInfLights / handleZero, direct.cpp:415-464
					const bool mono		  = rayFlags & PRB_RAY_MONOCHROME;
					const Blob heroFactor = mono ? heroOnly() : blob(1);
					if (S.hasInfLight && st.do_direct) {
						float denom_mis = 0;
						Blob radiance	= blob(0);
						for (uint32_t i = 0; i < S.nLights; ++i) {
							const prb_light& l = S.lights[i];
							if (!isInfLight(l) || isDeltaLight(l))
								continue;
							Blob rad;
							float pdfS;
							infLightEval(S, l, D, depth, wvl, rad, pdfS);
							const float pdf_S = pdfS * l.select_pdf;
							radiance		  = radiance + rad;
							denom_mis += bsum(misTermB(power, PrevPathPDF * pdf_S));
						}
						Blob mis;
						if (!st.do_nee || LastWasDelta) {
							mis = heroFactor / (WavelengthPDF * bsum(heroFactor));
						} else {
							const float denom = bsum(misTermB(power, PathPDF)) + denom_mis;
							mis				  = (heroFactor * misTerm(power, PathPDF[0])) / (misTermB(power, WavelengthPDF) * denom);
						}
						if (fragmentXYZ(S, mis, Throughput, radiance, rayFlags, groupMono, wvl, xyz)) {
							acc[0] += xyz[0];
							acc[1] += xyz[1];
							acc[2] += xyz[2];
						}
					}
				}
			} else {
				// ---------------- hit: makeIP (RenderTileSession::traceSingleRay :80-101, IntersectionPoint::setForSurface :61-75)
				const float t = W.hitT[slot];
				const V3 P	  = O + t * D;
				GeomPoint g;
				provideGeometryPoint(S, hraw.x, hraw.y, __uint_as_float(hraw.z), __uint_as_float(hraw.w), P, g);
				const float depth2 = norm2(O - P);
				const float NdotV  = dot(D, g.N);
				// ---------------- handleCameraVertex, direct.cpp:73-105
				++sEntity;
				++sDepth;
				if (depth == 0) { // pushSPFragment -> commitShadingPoints, LocalFrameOutputDevice.cpp:252-302
					W.sampleCount[pix] += 1;
					if (W.aov) {
						float* a = W.aov + 10 * (size_t)pix;
						a[0] += g.N.x;
						a[1] += g.N.y;
						a[2] += g.N.z;
						a[3] += P.x;
						a[4] += P.y;
						a[5] += P.z;
						a[6] += g.u;
						a[7] += g.v;
						a[8] += sqrtf(depth2);
						a[9] += (float)g.entity;
					}
				}
				const bool hasEmission = g.emission != PRB_INVALID_ID;
				bool cont			   = true;
				if (st.do_direct && hasEmission) {
					// ------------ handleDirectHit, direct.cpp:355-412
					if (g.emission < S.nEmissions) {
						const float cosC = -NdotV;
						if (!(fabsf(cosC) <= PR_EPSILON)) {
							const bool hitFromBehind = cosC < 0.0f;
							const Blob radiance		 = hitFromBehind ? blob(0) : evalNode(S, S.emissions[g.emission].radiance_node, wvl, g.u, g.v);
							const bool mono			 = rayFlags & PRB_RAY_MONOCHROME;
							const Blob heroFactor	 = mono ? heroOnly() : blob(1);
							Blob mis;
							if (!st.do_nee || hitFromBehind || LastWasDelta) {
								mis = heroFactor / (WavelengthPDF * bsum(heroFactor));
							} else {
								const prb_entity& en = S.entities[g.entity];
								const float selProb	 = en.light_id != PRB_INVALID_ID ? S.lights[en.light_id].select_pdf : 0.0f;
								float posPDF		 = 0;
								if (en.light_id != PRB_INVALID_ID) {
									const float4 lp = W.lastPos[slot];
									posPDF			= entityPositionPDF(S, g.entity, P, mk(lp.x, lp.y, lp.z));
									posPDF			= posPDF * depth2 / fabsf(cosC); // IS::toSolidAngle
								}
								const float posPDF_S = posPDF * selProb;
								const float denom	 = bsum(misTermB(power, PrevPathPDF * posPDF_S)) + bsum(misTermB(power, PathPDF));
								mis					 = (heroFactor * misTerm(power, PathPDF[0])) / (misTermB(power, WavelengthPDF) * denom);
							}
							if (fragmentXYZ(S, mis, Throughput, radiance, rayFlags, groupMono, wvl, xyz)) {
								acc[0] += xyz[0];
								acc[1] += xyz[1];
								acc[2] += xyz[2];
							}
						}
					}
					if (!st.emissive_scatter)
						cont = false;
				}
				const uint32_t matID = g.material;
				if (matID >= S.nMaterials)
					cont = false;
				if (cont) {
					Rng rnd{ W.rng[pix] };
					const bool onlyDelta = S.materials[matID].flags & PRB_MATF_ONLY_DELTA;
					MatCtx mc;
					mc.V		= toTangentSpace(g.N, g.Nx, g.Ny, -D);
					mc.wvl		= wvl;
					mc.u		= g.u;
					mc.v		= g.v;
					mc.rayFlags = rayFlags;
					if (st.do_nee && !onlyDelta && !hasEmission && S.nLights > 0) {
						// -------- handleNEE, direct.cpp:233-352
						const float usel  = rnd.getFloat();
						const int lightID = cdfSearch(S.lightCDF, (int)S.nLights + 1, usel);
						const float selPdf = S.lightCDF[lightID + 1] - S.lightCDF[lightID];
						const prb_light& light = S.lights[lightID];
						LightSample ls;
						sampleLight(S, light, P, wvl, rnd, ls);
						const float sqrD = norm2(ls.lightPos - P);
						const V3 L		 = ls.outgoing;
						const float cosC = fabsf(dot(L, g.N));
						const float cosL = fabsf(ls.cosLight);
						if (cosC * cosL > 1e-5f && sqrD > 1e-5f) { // GEOMETRY_EPS / DISTANCE_EPS
							mc.L = toTangentSpace(g.N, g.Nx, g.Ny, L);
							MatEval mout;
							materialEval<MATERIALS>(S, matID, mc, mout);
							if (!(mout.flags & MSF_Delta)) {
								const bool rayMono		 = rayFlags & PRB_RAY_MONOCHROME;
								const bool bsdfMono		 = rayMono; // a non-delta eval result is never hero collapsing
								const Blob rayHeroFactor = rayMono ? heroOnly() : blob(1);
								const Blob heroFactor	 = bsdfMono ? heroOnly() : blob(1);
								const Blob bsdfWvlPdfS	 = mout.pdf * heroFactor;
								if (!allLE(bsdfWvlPdfS, 1e-6f)) { // PDF_EPS
									const Blob connectionW = ls.radiance * mout.weight;
									const bool worthACheck = !blobIsZero(connectionW, PR_EPSILON);
									float lightPdfS		   = ls.infinite ? ls.dirPDF_S : ls.posPDF * sqrD / cosL;
									lightPdfS *= selPdf;
									if (ls.delta) // light->hasDeltaDistribution(), direct.cpp:288-289
										lightPdfS = 1;
									const bool normalPdf = !(isnan(lightPdfS) || isinf(lightPdfS) || lightPdfS == 0.0f || fabsf(lightPdfS) < 1.17549435e-38f);
									if (ls.delta || (normalPdf && !(lightPdfS <= 1e-6f))) {
										const Blob lightPdfS2 = rayHeroFactor * lightPdfS;
										if (!allLE(lightPdfS2, 1e-6f)) {
											Blob mis;
											if (st.do_direct && !LastWasEmissive) {
												const float cameraRoulette = rrProbability(S, depth + 1, false);
												const Blob bsdfPdfS		   = bsdfWvlPdfS * cameraRoulette;
												const float denom = bsum(misTermB(power, PathPDF * lightPdfS2)) + bsum(misTermB(power, PathPDF * bsdfPdfS));
												mis = ls.delta ? heroFactor / bsum(heroFactor)
															   : blob(misTerm(power, PathPDF[0] * lightPdfS2[0])) / ((heroFactor * denom) * misTermB(power, WavelengthPDF));
											} else {
												mis = heroFactor / (WavelengthPDF * bsum(heroFactor));
											}
											const float distance = ls.infinite ? PRB_INF : sqrtf(sqrD);
											// shadow ray: cameraIP.nextRay(L, Shadow, SHADOW_RAY_MIN, distance), IntersectionPoint.h:116-124
											const V3 oN		= dot(L, g.N) < 0 ? -g.N : g.N;
											const V3 sO		= safePosition(P, L, oN);
											const uint32_t shadowFlags = rayFlags | PRB_RAY_SHADOW;
											if (ls.infinite)
												++sBg;
											else
												++sEntity;
											if (worthACheck) {
												++sShadow;
												const Blob contrib = connectionW / lightPdfS2[0];
												if (fragmentXYZ(S, mis, Throughput, contrib, shadowFlags, groupMono, wvl, xyz)) {
													// Scene::traceShadowRay: tnear = MinT (1e-4), tfar = distance - 0.001, Scene.cpp:266-280
													W.shO[slot]	  = make_float4(sO.x, sO.y, sO.z, 0.0001f);
													W.shD[slot]	  = make_float4(L.x, L.y, L.z, distance - 0.001f);
													W.shXYZ[slot] = make_float4(xyz[0], xyz[1], xyz[2], 0);
													pushShadow	  = true;
												}
											}
										}
									}
								}
							}
						}
					}
					LastWasEmissive = hasEmission;
					// -------- handleScattering, direct.cpp:170-230
					const float scatProb = rrProbability(S, depth + 1, onlyDelta);
					bool scatter		 = scatProb > PR_EPSILON;
					// RussianRoulette::check: one draw only when 0 < p < 1.  Written branch-free on purpose: the nested form
					// `if (p < 1) { if (draw > p) scatter = false; }` is folded to "always kill" by nvcc 12.9's NVVM here
					// (seen in the PTX; caught by the oracle parity test), this form is compiled faithfully.
					float rrDraw = 0.0f;
					if (scatter && scatProb < 1.0f)
						rrDraw = rnd.getFloat();
					scatter = scatter && !(rrDraw > scatProb);
					if (scatter) {
						MatSample sout;
						mc.L = mk(0, 0, 0);
						materialSample<MATERIALS>(S, matID, mc, rnd, sout);
						const V3 L	 = normalized(fromTangentSpace(g.N, g.Nx, g.Ny, sout.L)); // MaterialSampleOutput::globalL
						LastWasDelta = sout.isDelta();
						PrevPathPDF	 = PathPDF;
						PathPDF		 = PathPDF * (sout.pdf * scatProb);
						if (!allLE(PathPDF, 1e-6f)) {
							Throughput = Throughput * sout.weight;
							if (sout.isHeroCollapsing()) {
								Throughput = Throughput * heroOnly();
								PathPDF	   = PathPDF * heroOnly();
							}
							if (!blobIsZero(Throughput, PR_EPSILON)) {
								uint32_t nf = rayFlags | PRB_RAY_BOUNCE;
								if (sout.isHeroCollapsing())
									nf |= PRB_RAY_MONOCHROME;
								const uint32_t nd = depth + 1;
								if (nd < st.max_ray_depth) { // Walker::traverse loop bound, vcm/Walker.h:26
									const V3 oN		 = dot(L, g.N) < 0 ? -g.N : g.N;
									const V3 nO		 = safePosition(P, L, oN);
									W.rayO[slot]	 = make_float4(nO.x, nO.y, nO.z, 0.0001f); // BOUNCE_RAY_MIN
									W.rayD[slot]	 = make_float4(L.x, L.y, L.z, PRB_INF);
									W.thr[slot]		 = tof4(Throughput);
									W.pathPDF[slot]	 = tof4(PathPDF);
									W.prevPDF[slot]	 = tof4(PrevPathPDF);
									W.lastPos[slot]	 = make_float4(P.x, P.y, P.z, 0);
									W.flagsDepth[slot] = nd | (nf << FD_FLAGS_SHIFT) | (LastWasDelta ? FD_LAST_DELTA : 0) | (LastWasEmissive ? FD_LAST_EMISSIVE : 0);
									alive = true;
									++sBounce;
									if (nf & PRB_RAY_MONOCHROME)
										++sMono;
								}
							}
						}
					}
					W.rng[pix] = rnd.s;
				}
			}
			uint32_t nst = pushShadow ? SF_SHADOW : 0u;
			if (alive) {
				W.iterXYZ[slot] = make_float4(acc[0], acc[1], acc[2], 0);
				nst |= SF_ACTIVE;
			} else {
				// the path ended: fold the sample into the film -- now, or in the next k_trace when its last NEE shadow ray
				// is still pending -- and refill the slot with the pixel's next camera sample
				const uint32_t iterCount = W.iter[slot];
				if (pushShadow) {
					W.prevAcc[slot] = make_float4(acc[0], acc[1], acc[2], __uint_as_float(iterCount));
					nst |= SF_FINALIZE;
				} else {
					foldSampleIntoFilm(W, pix, acc[0], acc[1], acc[2], iterCount);
				}
				W.iterXYZ[slot] = make_float4(0, 0, 0, 0);
				regen			= true;
			}
			W.state[slot] = nst;
		}
		// ---- camera-sample regeneration, compacted over the block: only the paths that ended in this pass (about one in
		// three on the Cornell box) need a new camera sample; run per lane it executed at 6 of 32 lanes and took 19 % of the
		// issue slots (profiles/r01_ncu_c2_v3.txt).  The ended slots are listed in shared memory and regenerated by the
		// first threads of the block, full warps at a time.
		if (regen)
			regenList[atomicAdd(&regenCount[round], 1u)] = (uint16_t)(slot - base);
		__syncthreads();
		if (threadIdx.x < regenCount[round]) {
			const uint32_t rs = base + regenList[threadIdx.x];
			if (startNextSample(S, W, rs, W.pixel[rs])) {
				W.state[rs] |= SF_ACTIVE;
				++sSamples;
			} else {
				atomicAdd(W.counters + CNT_RETIRED, 1u);
			}
		}
		__syncthreads(); // regenList is reused by the next pass
	}
	if (blockIdx.x == 0 && threadIdx.x == 0)
		W.counters[CNT_WORK] = 0; // work counter of the next k_trace (stream order: this kernel runs after k_trace finished)
	statAdd(W.stats, ST_PIXEL_SAMPLE, sSamples);
	statAdd(W.stats, ST_CAMERA_RAY, sSamples);
	statAdd(W.stats, ST_PRIMARY, sSamples);
	statAdd(W.stats, ST_ENTITY_HIT, sEntity);
	statAdd(W.stats, ST_BG_HIT, sBg);
	statAdd(W.stats, ST_CAMERA_DEPTH, sDepth);
	statAdd(W.stats, ST_SHADOW, sShadow);
	statAdd(W.stats, ST_BOUNCE, sBounce);
	statAdd(W.stats, ST_CAMERA_RAY, sBounce);
	statAdd(W.stats, ST_MONO, sMono);
}

// ------------------------------------------------------------------ stream tracing (prb_trace_closest / _any)
// persistent-thread ray-stream kernels: `counter` must be zero at launch
template <bool ANY>
PRB_DEV void traceStream(const DScene& S, const float* ox, const float* oy, const float* oz, const float* dx, const float* dy, const float* dz, const float* tmin,
						 const float* tmax, uint32_t n, uint32_t* counter, uint32_t* ent, uint32_t* prim, float* u, float* v, float* t, uint8_t* occluded)
{
	Trav tr;
	uint2 stack[BVH_STACK_ALLOC];
	tr.idle();
	uint32_t i	= 0;
	bool active = false, pool = true;
	float t1	= 0;
	for (;;) {
		{
			const bool need	 = !active && pool;
			const uint32_t k = fetchWork(counter, need);
			if (need) {
				if (k < n) {
					i			   = k;
					const float t0 = tmin ? tmin[i] : 0.0001f;
					t1			   = tmax ? tmax[i] : PRB_INF;
					tr.begin(S, mk(ox[i], oy[i], oz[i]), mk(dx[i], dy[i], dz[i]), t0, t1, ANY, stack);
					active = true;
				} else {
					pool = false;
				}
			}
		}
		if (__ballot_sync(0xFFFFFFFFu, active) == 0)
			break;
		for (;;) { // warp-synchronous rounds
			if (tr.round(S, active, stack)) {
				const bool ok = tr.hit();
				if (ANY) {
					occluded[i] = ok ? 1 : 0;
				} else {
					ent[i]	= ok ? tr.best.entity : PRB_INVALID_ID;
					prim[i] = ok ? tr.best.prim : PRB_INVALID_ID;
					u[i]	= ok ? tr.best.u : 0.0f;
					v[i]	= ok ? tr.best.v : 0.0f;
					t[i]	= ok ? tr.best.t : t1;
				}
				active = false;
			}
			const unsigned am = __ballot_sync(0xFFFFFFFFu, active);
			if (am == 0 || (__popc(am) < REFILL_LANES && __any_sync(0xFFFFFFFFu, pool && !active)))
				break;
		}
	}
}
__global__ void __launch_bounds__(128) k_trace_closest(DScene S, const float* ox, const float* oy, const float* oz, const float* dx, const float* dy,
														const float* dz, const float* tmin, const float* tmax, uint32_t n, uint32_t* counter, uint32_t* ent,
														uint32_t* prim, float* u, float* v, float* t)
{
	traceStream<false>(S, ox, oy, oz, dx, dy, dz, tmin, tmax, n, counter, ent, prim, u, v, t, nullptr);
}
__global__ void __launch_bounds__(128) k_trace_any(DScene S, const float* ox, const float* oy, const float* oz, const float* dx, const float* dy,
													const float* dz, const float* tmin, const float* tmax, uint32_t n, uint32_t* counter, uint8_t* occluded)
{
	traceStream<true>(S, ox, oy, oz, dx, dy, dz, tmin, tmax, n, counter, nullptr, nullptr, nullptr, nullptr, nullptr, occluded);
}

// camera rays only (prb_generate_camera_rays): does not touch the RNG map
__global__ void k_camera_rays(const __grid_constant__ DScene S, const uint64_t* rng, const uint32_t* pixels, uint32_t n, uint32_t iteration, float* org, float* dir, float* wvl)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint32_t pix = pixels[i];
		Rng rnd{ rng[pix] };
		CameraSampleOut cs;
		const uint32_t fw = S.settings.film_width;
		constructCameraRay(S, pix % fw, pix / fw, iteration, rnd, cs);
		org[3 * i] = cs.origin.x, org[3 * i + 1] = cs.origin.y, org[3 * i + 2] = cs.origin.z;
		dir[3 * i] = cs.dir.x, dir[3 * i + 1] = cs.dir.y, dir[3 * i + 2] = cs.dir.z;
		for (int k = 0; k < 4; ++k)
			wvl[4 * i + k] = cs.wvl[k];
	}
}

// unit-level material calls (IMaterial::eval / ::sample)
__global__ void k_material_eval(const __grid_constant__ DScene S, const prb_material_query* q, uint32_t n, prb_material_result* out)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		MatCtx c;
		c.V = ld3(q[i].V);
		c.L = ld3(q[i].L);
		for (int k = 0; k < 4; ++k)
			c.wvl[k] = q[i].wavelength_nm[k];
		c.u		   = q[i].uv[0];
		c.v		   = q[i].uv[1];
		c.rayFlags = q[i].ray_flags;
		MatEval e;
		materialEval<SHADE_MATERIALS_COMBINED>(S, q[i].material_id, c, e);
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
__global__ void k_material_sample(const __grid_constant__ DScene S, const prb_material_query* q, uint32_t n, prb_material_result* out)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		MatCtx c;
		c.V = ld3(q[i].V);
		c.L = mk(0, 0, 0);
		for (int k = 0; k < 4; ++k)
			c.wvl[k] = q[i].wavelength_nm[k];
		c.u		   = q[i].uv[0];
		c.v		   = q[i].uv[1];
		c.rayFlags = q[i].ray_flags;
		Rng rnd{ q[i].rng_state };
		MatSample e;
		materialSample<SHADE_MATERIALS_COMBINED>(S, q[i].material_id, c, rnd, e);
		for (int k = 0; k < 4; ++k) {
			out[i].weight[k] = e.weight[k];
			out[i].pdf_s[k]	 = e.pdf[k];
		}
		out[i].L[0] = e.L.x, out[i].L[1] = e.L.y, out[i].L[2] = e.L.z;
		out[i].flags	 = e.flags;
		out[i].type		 = e.type;
		out[i].rng_state = rnd.s;
	}
}

// pixel filter post pass (FilterCache table, zero padded) and film export / import
__global__ void k_filter(const float* in, float* out, int W, int H, int r, const float* tab)
{
	const int n = W * H;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const int x = i % W, y = i / W;
		float s[3] = { 0, 0, 0 };
		const int dia = 2 * r + 1;
		// gather form of the reference's splat: out[p] = sum_q w(p - q) * in[q]; the table is symmetric
		for (int dy = -r; dy <= r; ++dy)
			for (int dx = -r; dx <= r; ++dx) {
				const int sx = x - dx, sy = y - dy;
				if (sx < 0 || sy < 0 || sx >= W || sy >= H)
					continue;
				const float w = tab[(dy + r) * dia + (dx + r)];
				if (!(w > PR_EPSILON))
					continue;
				for (int c = 0; c < 3; ++c)
					s[c] += w * in[3 * (sy * W + sx) + c];
			}
		out[3 * i] = s[0], out[3 * i + 1] = s[1], out[3 * i + 2] = s[2];
	}
}
__global__ void k_film_export(const float* mean, const uint32_t* count, float* dst, uint32_t n)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		dst[4 * i]	   = mean[3 * i];
		dst[4 * i + 1] = mean[3 * i + 1];
		dst[4 * i + 2] = mean[3 * i + 2];
		dst[4 * i + 3] = (float)count[i];
	}
}
__global__ void k_film_import(const float* src, float* mean, uint32_t* count, uint32_t n)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		mean[3 * i]		= src[4 * i];
		mean[3 * i + 1] = src[4 * i + 1];
		mean[3 * i + 2] = src[4 * i + 2];
		count[i]		= (uint32_t)src[4 * i + 3];
	}
}
} // namespace prb
