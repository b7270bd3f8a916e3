// sort.cu -- segmented, stable LSD radix sort of the per-effect (depth key, slot) pairs
// (one segment per effect, 8-bit digits, onesweep-style chained scan per digit) and the
// 4x vertex expansion in sorted order.
//
// Extension row A11/K4 of SURVEY.md: the reference sorts effects, not particles; the
// per-particle order follows the BillboardOrder idiom (src/client/BillboardSystem.cpp:41-51):
// depth descending (back to front), ties by ascending slot. The keys were written by
// k_integrate in ascending slot order as ~enc(depth), so a STABLE ascending sort gives
// exactly that order.
//
//   k_sort_hist    per tile: digit histograms of all four passes (keys read once)
//   k_sort_scan    per (segment, pass): exclusive scan of the 256 bins -> digit bases; marks the
//                  passes whose digit is constant over the segment (they are skipped)
//   k_sort_pass    x4: rank keys inside a 4096-key tile (warp match ranking, stable),
//                  chained look-back per digit across the tiles of the segment,
//                  stage the tile in shared memory in sorted order, coalesced scatter
//   k_expand_sorted  rank r -> slot -> 32-byte record -> 4 x 32 B at rank r (:607-665)

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "device_types.h"
#include "kernels.h"
#include "launch.h"

namespace spr {

#define FULL_MASK 0xffffffffu

#ifndef SPR_SORT_ITEMS
#define SPR_SORT_ITEMS 16
#endif
static const uint32_t kItemsPerThread = SPR_SORT_ITEMS;               // keys per thread of a pass CTA
static const uint32_t kSortThreads = kSortTileKeys / kItemsPerThread; // 256 threads for 4096-key tiles, 512 for 8192
static const uint32_t kSortWarps = kSortThreads / 32;
static const uint32_t kWarpKeys = 32 * kItemsPerThread;               // 512 keys per warp
static const uint32_t kHistItems = kSortTileKeys / 256;               // keys per thread of the (256-thread) histogram CTA
static const uint32_t kExpandPerTile = kSortTileKeys / 1024;          // k_expand_sorted CTAs per sort tile
static const uint32_t kLook = 8;                                      // status words fetched per look-back batch
static const size_t kPassSmemBytes = (2 * kSortTileKeys + kSortWarps * 256) * sizeof(uint32_t); // keys, values, per-warp digit counts
#ifndef SPR_SORT_MIN_CTAS
#define SPR_SORT_MIN_CTAS (1024 / kSortThreads)
#endif
static_assert(kSortThreads >= 256 && kSortThreads <= 1024 && kHistItems <= 32, "sort tile: 4096 to 16384 keys");
static_assert(kSortWarps * 512 <= 2 * kSortTileKeys, "the per-warp peer-mask tables alias the key / value staging buffers");

__global__ void __launch_bounds__(256) k_sort_hist(PoolPtrs pool, TablePtrs tab, const uint32_t *segEffects, const SortTile *tiles, uint32_t *hist)
{
	pdl_enter();
	__shared__ uint32_t sHist[4 * 256];
	SortTile st = tiles[blockIdx.x];
	uint32_t g = segEffects[st.effect];
	const EffectState &S = tab.state[g];
	uint32_t slots = S.slotCount[S.parity];
	uint32_t begin = st.tile * kSortTileKeys;
	if (begin >= slots) return;
	uint32_t count = min(kSortTileKeys, slots - begin);
	for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) sHist[i] = 0;
	__syncthreads();
	const uint32_t segBase = tab.params[g].slotBase;
	const uint32_t *keys = pool.keys + segBase + begin;
	const uint32_t *mask = pool.aliveMask + ((segBase + begin) >> 5);
	const uint32_t keyMin = effect_key_min(S); // digits are taken from key - keyMin: same order, fewer significant bits
	// All loads of the tile first and with no branch between them -- one memory round trip per CTA instead of one per
	// key (alive word -> branch -> key was two dependent loads, sixteen times over: the kernel took the same 42 us for
	// 1M and for 10M keys). Key i = tid + 256 * it sits in alive word warp + 8 * it, bit lane: lane l < 16 fetches the
	// word of iteration l (lane l < kHistItems), a dead slot's key is stale but mapped.
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint32_t myWord = 0;
	if (lane < kHistItems && (warp + 8 * lane) * 32 < count) myWord = mask[warp + 8 * lane];
	uint32_t key[kHistItems];
#pragma unroll
	for (uint32_t it = 0; it < kHistItems; it++) {
		const uint32_t i = threadIdx.x + it * 256u;
		key[it] = i < count ? keys[i] : 0u;
	}
#pragma unroll
	for (uint32_t it = 0; it < kHistItems; it++) {
		const uint32_t i = threadIdx.x + it * 256u;
		const uint32_t word = __shfl_sync(FULL_MASK, myWord, it);
		if (i >= count || !((word >> lane) & 1u)) continue;
		const uint32_t k = key[it] - keyMin;
		// (one count per warp for the top digits, which are the same for most keys of an effect, was measured twice:
		// slower -- same-address shared-memory atomics are not what this kernel waits for)
		atomicAdd(&sHist[0 * 256 + (k & 255u)], 1u);
		atomicAdd(&sHist[1 * 256 + ((k >> 8) & 255u)], 1u);
		atomicAdd(&sHist[2 * 256 + ((k >> 16) & 255u)], 1u);
		atomicAdd(&sHist[3 * 256 + (k >> 24)], 1u);
	}
	__syncthreads();
	uint32_t *dst = hist + (size_t)st.effect * 1024;
	for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) {
		uint32_t v = sHist[i];
		if (v) atomicAdd(&dst[i], v);
	}
}

// one warp per (segment, pass). passMask[segment] (zeroed with the histograms) gets bit p when pass p has
// to run: pass 0 always (its scatter is the compaction), a later pass only when its digit is not the same
// for every key of the segment. The digits are those of key - keyMin (EffectState::keyMin, the smallest key of
// the step: same order, fewer significant bits), and the depths of one effect are floats from a narrow range
// -- under 2^24 distinct values unless the effect spans more than about two octaves of squared distance --,
// so the top digit is normally 0 everywhere and that pass would move every pair to where it already is.
__global__ void k_sort_scan(TablePtrs tab, const uint32_t *segEffects, uint32_t *hist, uint32_t *passMask, uint32_t numRows)
{
	pdl_enter();
	uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (row >= numRows) return;
	uint32_t *h = hist + (size_t)row * 256;
	uint32_t v[8], sum = 0, big = 0;
#pragma unroll
	for (int i = 0; i < 8; i++) { v[i] = h[lane * 8 + i]; sum += v[i]; big = max(big, v[i]); }
	big = __reduce_max_sync(FULL_MASK, big);
	uint32_t incl = sum;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(FULL_MASK, incl, d); if (lane >= d) incl += o; }
	uint32_t run = incl - sum;
#pragma unroll
	for (int i = 0; i < 8; i++) { h[lane * 8 + i] = run; run += v[i]; }
	// the bins of any pass add up to the number of live particles (gpuParticles.size / 4, :670)
	if ((row & 3u) == 0 && lane == 31) tab.state[segEffects[row >> 2]].aliveCount = run;
	if (lane == 31 && ((row & 3u) == 0 || big != run)) atomicOr(&passMask[row >> 2], 1u << (row & 3u));
}

__device__ inline uint32_t ld_relaxed_u32(const uint32_t *p)
{
	uint32_t v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ inline void st_relaxed_u32(uint32_t *p, uint32_t v)
{
	asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// One radix pass over digit `shift/8`. lookback: [numTiles][256] status words of THIS pass,
// zero-initialised; word = [31:30] flag (1 aggregate, 2 inclusive prefix), [29:0] count.
// The pairs ping-pong between buffer 0 (PoolPtrs::keys/vals) and buffer 1 (the Alt pair): a segment that has
// executed e passes so far has its pairs in buffer e & 1.
// The LAST executed pass of a segment (the highest bit of its pass mask) has no reader for its keys: it writes only
// the sorted slots.
#ifdef SPR_SORT_PROFILE
__device__ unsigned long long g_sortProf[16];
#define PROF_T(i) do { if (threadIdx.x == 0) { long long t_ = clock64(); atomicAdd(&g_sortProf[i], (unsigned long long)(t_ - profT)); profT = t_; } } while (0)
#else
#define PROF_T(i) do { } while (0)
#endif

template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads, SPR_SORT_MIN_CTAS) k_sort_pass(PoolPtrs pool, TablePtrs tab, const uint32_t *segEffects, const SortTile *tiles,
	uint32_t *keys0, uint32_t *vals0, uint32_t *keys1, uint32_t *vals1,
	const uint32_t *hist, const uint32_t *passMask, uint32_t *lookback, uint32_t *ticket, uint32_t pass)
{
	pdl_enter();
	extern __shared__ __align__(16) uint32_t sPassDyn[];
	uint32_t *sKeys = sPassDyn;                         // [kSortTileKeys]
	uint32_t *sVals = sPassDyn + kSortTileKeys;         // [kSortTileKeys]
	uint32_t *sCnt = sPassDyn + 2 * kSortTileKeys;      // [kSortWarps][256] per-warp digit counts -> staging position of the warp's first key of each digit
	__shared__ int32_t sGlobal[256];            // global position of local index j with digit d: sGlobal[d] + j
	__shared__ uint32_t sWarpSum[8];
	__shared__ uint32_t sTileCount, sTicket;

	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#if SPR_STRICT_TILE_TICKETS
	// The CTA's place in the tile table is a ticket, taken in the order the CTAs start (not the block index): the
	// chained look-back below then only ever waits for tiles whose CTAs are already running.
	if (tid == 0) sTicket = atomicAdd(ticket, 1u);
	__syncthreads();
	const uint32_t tileSlot = sTicket;
#else
	// Tile = block index: the look-back below waits only for lower block indices of this grid, which the hardware
	// dispatched before this CTA (forward-progress argument: device_types.h, SPR_STRICT_TILE_TICKETS).
	const uint32_t tileSlot = blockIdx.x;
	(void)ticket; (void)sTicket;
#endif
	const SortTile st = tiles[tileSlot];
	const uint32_t g = segEffects[st.effect];
	const uint32_t runMask = passMask[st.effect] | 1u; // pass 0 always runs: its scatter is the compaction
	if (!FIRST && !((runMask >> pass) & 1u)) return; // constant digit: nothing would move
	const bool isLast = (runMask >> (pass + 1u)) == 0u;
	const bool srcIs1 = !FIRST && (__popc(runMask & ((1u << pass) - 1u)) & 1u);
	const uint32_t *keysIn = srcIs1 ? keys1 : keys0, *valsIn = srcIs1 ? vals1 : vals0;
	uint32_t *keysOut = srcIs1 ? keys0 : keys1, *valsOut = srcIs1 ? vals0 : vals1;
	// the first pass reads the keys in SLOT space (holes masked out by the alive ballots) and its
	// scatter compacts them; later passes run on the dense (key, slot) pairs
	const uint32_t alive = FIRST ? tab.state[g].slotCount[tab.state[g].parity] : tab.state[g].aliveCount;
	const uint32_t begin = st.tile * kSortTileKeys;
	if (begin >= alive) return;
	const uint32_t count = min(kSortTileKeys, alive - begin);
	const uint32_t segBase = tab.params[g].slotBase;
	const uint32_t shift = pass * 8;
	const uint32_t keyMin = effect_key_min(tab.state[g]);
	const uint32_t *mask = pool.aliveMask + ((segBase + begin) >> 5) + warp * (kWarpKeys / 32);

#ifdef SPR_SORT_PROFILE
	long long profT = clock64();
#endif
	// per-warp digit counts and the two per-warp peer-mask tables (aliased onto sKeys / sVals, idle until staging)
	for (uint32_t i = tid; i < kSortWarps * 256; i += kSortThreads) { sCnt[i] = 0; sPassDyn[i] = 0; sPassDyn[i + kSortWarps * 256] = 0; }
	__syncthreads();
	PROF_T(0);

	// ---- load keys (warp-striped: warp w owns keys [w*512, w*512+512) of the tile) and rank them.
	// Register diet: ranks and staging positions are 16-bit and packed in pairs, the
	// slot values are only fetched once the keys have been staged.
	uint32_t key[kItemsPerThread], rank2[kItemsPerThread / 2];
	uint32_t okBits = 0;
	const uint32_t *kin = keysIn + segBase + begin;
	const uint32_t *vin = valsIn + segBase + begin;
	// every load of the tile is issued before anything depends on one: the first pass fetches the warp's 16 alive
	// words with one coalesced load (lane i holds the word of item i) and its keys whether alive or not (a dead slot's
	// key is stale, not unmapped)
	uint32_t myWord = 0;
	if (FIRST && lane < kItemsPerThread && warp * kWarpKeys + lane * 32 < count) myWord = mask[lane];
#pragma unroll
	for (uint32_t i = 0; i < kItemsPerThread; i++) {
		uint32_t idx = warp * kWarpKeys + i * 32 + lane;
		key[i] = idx < count ? kin[idx] : 0xffffffffu;
	}
#pragma unroll
	for (uint32_t i = 0; i < kItemsPerThread; i++) {
		uint32_t idx = warp * kWarpKeys + i * 32 + lane;
		bool ok = idx < count;
		if (FIRST) {
			const uint32_t word = __shfl_sync(FULL_MASK, myWord, i);
			ok = ok && ((word >> lane) & 1u);
			if (!ok) key[i] = 0xffffffffu;
		}
		okBits |= (ok ? 1u : 0u) << i;
	}
#ifdef SPR_SORT_PROFILE
	if (key[0] == 0x12345678u && key[15] == 0x9abcdef0u) sWarpSum[0] = 1; // force the loads to have landed
#endif
	PROF_T(1);
	uint32_t *cnt = sCnt + warp * 256;
	const uint32_t ltMask = (1u << lane) - 1u;
	// Peers of a lane = lanes of the warp with the same digit, found with one shared-memory atomicOr
	// per lane into a per-warp mask table (measured 14 % faster per pass than MATCH.ANY, whose cost
	// grows with the number of distinct digits in the warp). Two tables used by alternate rounds:
	// the leader clears its entry after reading it, and two __syncwarp()s pass before its next use.
	// (two keys per round through four tables -- half the warp barriers -- was measured 4 % slower:
	// profiles/r02_sort_experiments.md)
	uint32_t *match0 = sPassDyn + warp * 512, *match1 = match0 + 256;
#pragma unroll
	for (uint32_t i = 0; i < kItemsPerThread; i++) {
		uint32_t *match = (i & 1u) ? match1 : match0;
		bool ok = (okBits >> i) & 1u;
		uint32_t d = ((key[i] - keyMin) >> shift) & 255u;
		if (ok) atomicOr(&match[d], 1u << lane);
		__syncwarp();
		uint32_t peers = ok ? match[d] : 0u;
		uint32_t before = ok ? cnt[d] : 0u;
		__syncwarp();
		uint32_t rk = before + __popc(peers & ltMask);
		if (ok && (peers & ltMask) == 0) { cnt[d] = before + __popc(peers); match[d] = 0; }
		if (i & 1u) rank2[i >> 1] |= rk << 16; else rank2[i >> 1] = rk;
	}
	PROF_T(2);
	__syncthreads();
	PROF_T(3);

	// ---- thread d < 256: per-warp exclusive offsets of digit d, tile count, chained look-back
	uint32_t total = 0, excl = 0, incl = 0;
	if (tid < 256) {
#pragma unroll
		for (uint32_t w = 0; w < kSortWarps; w++) { uint32_t c = sCnt[w * 256 + tid]; sCnt[w * 256 + tid] = total; total += c; }
		uint32_t *lb = lookback + (size_t)tileSlot * 256 + tid;
		if (st.tile == 0) {
			st_relaxed_u32(lb, (2u << 30) | total);
		} else {
			st_relaxed_u32(lb, (1u << 30) | total);
			// walk back over the predecessors' status words of this digit, kLook at a time: the loads of one batch are
			// independent, so a long walk costs one L2 round trip per batch instead of one per tile
			const uint32_t *p = lb - 256; // tiles of one segment are consecutive in the tile table
			bool found = false;
			for (uint32_t t = st.tile; t > 0 && !found; ) {
				const uint32_t n = t < kLook ? t : kLook;
				uint32_t v[kLook];
#pragma unroll
				for (uint32_t i = 0; i < kLook; i++) v[i] = i < n ? ld_relaxed_u32(p - i * 256) : (2u << 30);
#pragma unroll
				for (uint32_t i = 0; i < kLook; i++) {
					if (found) break;
					// (a predecessor tile is always already running; trap instead of hanging if that ever breaks)
					for (uint32_t spins = 0; (v[i] >> 30) == 0; ) { v[i] = ld_relaxed_u32(p - i * 256); if (++spins > (1u << 24)) __trap(); }
					excl += v[i] & 0x3fffffffu;
					found = (v[i] >> 30) == 2u;
				}
				p -= kLook * 256; t -= n;
			}
			st_relaxed_u32(lb, (2u << 30) | (excl + total));
		}
		PROF_T(4);
		// exclusive scan of the 256 digit counts of the tile
		incl = total;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(FULL_MASK, incl, d); if (lane >= d) incl += o; }
		if (lane == 31) sWarpSum[warp] = incl;
	}
	__syncthreads();
	if (tid < 256) {
		uint32_t warpBase = 0;
#pragma unroll
		for (uint32_t w = 0; w < 8; w++) if (w < warp) warpBase += sWarpSum[w];
		const uint32_t tileExcl = warpBase + incl - total;
		// staging position of warp w's first key of digit d: one table read per key when the tile is staged
#pragma unroll
		for (uint32_t w = 0; w < kSortWarps; w++) sCnt[w * 256 + tid] += tileExcl;
		if (tid == 255) sTileCount = tileExcl + total;
		sGlobal[tid] = (int32_t)(hist[((size_t)st.effect * 4 + pass) * 256 + tid] + excl) - (int32_t)tileExcl;
	}
	__syncthreads();
	PROF_T(5);

	// ---- stage the tile in sorted order: keys first, then the values behind them
#pragma unroll
	for (uint32_t i = 0; i < kItemsPerThread; i++) {
		if ((okBits >> i) & 1u) {
			uint32_t d = ((key[i] - keyMin) >> shift) & 255u;
			uint32_t rk = (i & 1u) ? (rank2[i >> 1] >> 16) : (rank2[i >> 1] & 0xffffu);
			uint32_t pos = sCnt[warp * 256 + d] + rk;
			sKeys[pos] = key[i];
			if (i & 1u) rank2[i >> 1] = (rank2[i >> 1] & 0xffffu) | (pos << 16); else rank2[i >> 1] = (rank2[i >> 1] & 0xffff0000u) | pos;
		}
	}
	{
		uint32_t val[kItemsPerThread];
#pragma unroll
		for (uint32_t i = 0; i < kItemsPerThread; i++) {
			uint32_t idx = warp * kWarpKeys + i * 32 + lane;
			if (FIRST) val[i] = begin + idx;
			else val[i] = ((okBits >> i) & 1u) ? vin[idx] : 0u;
		}
#pragma unroll
		for (uint32_t i = 0; i < kItemsPerThread; i++) {
			if ((okBits >> i) & 1u) sVals[(i & 1u) ? (rank2[i >> 1] >> 16) : (rank2[i >> 1] & 0xffffu)] = val[i];
		}
	}
	PROF_T(6);
	__syncthreads();
	PROF_T(7);

	// ---- scatter: consecutive threads write consecutive addresses inside each digit run
	const uint32_t tileCount = sTileCount;
	uint32_t *vout = valsOut + segBase;
	if (!isLast) {
		uint32_t *kout = keysOut + segBase;
		for (uint32_t j = tid; j < tileCount; j += kSortThreads) {
			uint32_t k = sKeys[j];
			uint32_t d = ((k - keyMin) >> shift) & 255u;
			uint32_t dst = (uint32_t)(sGlobal[d] + (int32_t)j);
			kout[dst] = k;
			vout[dst] = sVals[j];
		}
	} else {
		for (uint32_t j = tid; j < tileCount; j += kSortThreads) {
			uint32_t d = ((sKeys[j] - keyMin) >> shift) & 255u;
			vout[(uint32_t)(sGlobal[d] + (int32_t)j)] = sVals[j];
		}
	}
	PROF_T(8);
#ifdef SPR_SORT_PROFILE
	if (threadIdx.x == 0) atomicAdd(&g_sortProf[15], 1ull);
#endif
}

#ifdef SPR_SORT_PROFILE
void sort_profile_dump()
{
	unsigned long long h[16];
	cudaDeviceSynchronize();
	cudaMemcpyFromSymbol(h, g_sortProf, sizeof(h));
	const char *names[9] = { "clear+sync", "key loads", "ranking", "sync after ranking", "offsets+lookback", "scan+tables+sync", "staging", "sync after staging", "scatter" };
	if (!h[15]) return;
	fprintf(stderr, "[sort profile] %llu CTAs; cycles per CTA (thread 0):", h[15]);
	for (int i = 0; i < 9; i++) fprintf(stderr, " %s %.0f;", names[i], (double)h[i] / (double)h[15]);
	fprintf(stderr, "\n");
	memset(h, 0, sizeof(h));
	cudaMemcpyToSymbol(g_sortProf, h, sizeof(h));
}
#endif

// ---------------------------------------------------------------------------
// Small segments: the whole sort of one effect in ONE CTA.
//
// An effect of up to kSmallSortMax slots (18432: the 16k-particle emitters of BASELINE configs[1] and everything
// smaller down to the 128-slot effects k_integrate_pass finishes itself) does not need the tile machinery: its
// (key, slot) pairs fit one CTA's registers and shared memory. One 1024-thread CTA per effect loads the keys in
// slot space under the alive ballots, runs the LSD passes over the digits of key - keyMin that are not all zero
// (same order as the tile path: stable, ascending), keeps keys in registers and permutes them through shared
// memory between passes (the 16-bit local slot index travels with them), and writes the sorted slot list to
// PoolPtrs::vals where k_expand_sorted reads it. No histogram kernel, no scan kernel, no look-back, no global
// ping-pong: for configs[1] this replaces 6 launches (hist, scan, 4 passes: ~85 us) by one of ~15 us.
// ---------------------------------------------------------------------------

static const uint32_t kSmallThreads = 1024;
static const uint32_t kSmallItems = kSmallSortMax / kSmallThreads; // 18 keys per thread at most
static const uint32_t kSmallWarps = kSmallThreads / 32;
// shared memory: keys (aliased by the per-warp peer-mask tables while ranking), 16-bit slots, per-warp digit counts
static const size_t kSmallSmemBytes = kSmallSortMax * 4 + kSmallSortMax * 2 + kSmallWarps * 256 * 4;
static_assert(kSmallItems * kSmallThreads == kSmallSortMax, "kSmallSortMax must be a multiple of the CTA size");
static_assert(kSmallWarps * 512 * 4 <= kSmallSortMax * 4, "the peer-mask tables alias the key buffer");
static_assert(kSmallSortMax <= 65536, "local slot indices are 16-bit");

// splitLog > 0 (a frame with few one-CTA effects, which would leave most SMs idle): 2 or 4 CTAs share an effect BY KEY RANGE.
// Every CTA of the group loads all keys of the effect, builds the same 256-bin histogram of their top 8 significant bits and
// cuts it into parts of about equal population; it keeps the keys of its own part -- the histogram's prefix tells where its
// piece of the sorted list starts --, compacts them in slot order and sorts them as below: fewer keys per thread, fewer
// bits (a pass less when the part's range drops below a byte boundary), no communication between the CTAs.
__global__ void __launch_bounds__(kSmallThreads, 1) k_sort_small(PoolPtrs pool, TablePtrs tab, const uint32_t *segEffects, uint32_t splitLog)
{
	pdl_launch_dependents();
	extern __shared__ __align__(16) uint8_t sRaw[];
	uint32_t *sKey = (uint32_t*)sRaw;                                  // [kSmallSortMax]
	uint16_t *sVal = (uint16_t*)(sRaw + kSmallSortMax * 4);            // [kSmallSortMax]
	uint32_t *sCnt = (uint32_t*)(sRaw + kSmallSortMax * 6);            // [kSmallWarps][256]
	__shared__ uint32_t sBase[256];     // exclusive scan of the digit totals
	__shared__ uint32_t sWarpSum[8];
	__shared__ uint32_t sMaxRel, sTotal;
	__shared__ uint32_t sPartWarp[kSmallWarps], sPivot[8];

	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t g = segEffects[blockIdx.x >> splitLog]; // host-uploaded list: read ahead of the wait
	const uint32_t part = blockIdx.x & ((1u << splitLog) - 1u);
	if (tid == 0) sMaxRel = 0;
	pdl_wait();
	EffectState &S = tab.state[g];
	const uint32_t slots = S.slotCount[S.parity];
	if (slots > kSmallSortMax) __trap(); // the host routes an effect here only when its slot bound fits
	const uint32_t segBase = tab.params[g].slotBase;
	const uint32_t keyMin = effect_key_min(S);
	// the CTA's keys are warp-striped: warp w owns [w * 32 * items, (w + 1) * 32 * items), item i of lane l is
	// w * 32 * items + i * 32 + l -- ascending in (warp, item, lane), which is the order ranks are given in
	uint32_t items = (slots + kSmallThreads - 1) / kSmallThreads;
	uint32_t warpBegin = warp * 32 * items;

	__syncthreads(); // sMaxRel = 0, written ahead of the wait

	uint32_t key[kSmallItems], rank2[kSmallItems / 2], val2[kSmallItems / 2];
	uint32_t okBits = 0, maxRel = 0;
	{
		// all loads first, with no branch between them (one memory round trip for the CTA, not one per item): lane l
		// fetches the alive word of the warp's item l, every lane its raw keys -- a dead slot's key is stale, not unmapped
		const uint32_t *kin = pool.keys + segBase;
		const uint32_t *mask = pool.aliveMask + (segBase >> 5); // segBase is a multiple of 128
		uint32_t myWord = 0;
		if (lane < items && warpBegin + lane * 32 < slots) myWord = mask[(warpBegin >> 5) + lane];
#pragma unroll
		for (uint32_t i = 0; i < kSmallItems; i++) {
			if (i >= items) break;
			const uint32_t idx = warpBegin + i * 32 + lane;
			key[i] = idx < slots ? kin[idx] : 0u;
		}
#pragma unroll
		for (uint32_t i = 0; i < kSmallItems; i++) {
			if (i >= items) break;
			const uint32_t idx = warpBegin + i * 32 + lane;
			const uint32_t word = __shfl_sync(FULL_MASK, myWord, i);       // bit `lane` of the word is slot idx
			const bool ok = idx < slots && ((word >> lane) & 1u);
			key[i] = ok ? key[i] - keyMin : 0u;
			if (ok) { okBits |= 1u << i; maxRel = max(maxRel, key[i]); }
			const uint32_t v = idx & 0xffffu;
			if (i & 1u) val2[i >> 1] |= v << 16; else val2[i >> 1] = v;
		}
	}
	maxRel = __reduce_max_sync(FULL_MASK, maxRel);
	if (lane == 0 && maxRel) atomicMax(&sMaxRel, maxRel);
	__syncthreads();
	maxRel = sMaxRel;
	uint32_t bits = maxRel ? 32u - (uint32_t)__clz(maxRel) : 0u;
	uint32_t outOffset = 0, totalAlive = 0;
	if (splitLog) {
		// ---- the parts: key ranges of about equal POPULATION, cut at boundaries of the 256 bins of the top 8 significant
		// bits (every CTA of the group computes the same histogram, hence the same cuts, from the same keys)
		const uint32_t numParts = 1u << splitLog;
		const uint32_t hshift = bits > 8u ? bits - 8u : 0u;
		const uint32_t ltMaskS = (1u << lane) - 1u;
		if (tid < 256) sBase[tid] = 0;
		if (tid < 8) sPivot[tid] = 256;
		__syncthreads();
#pragma unroll
		for (uint32_t i = 0; i < kSmallItems; i++) {
			if (i >= items) break;
			if ((okBits >> i) & 1u) atomicAdd(&sBase[key[i] >> hshift], 1u);
		}
		__syncthreads();
		uint32_t cntD = 0, inclD = 0;
		if (tid < 256) { // inclusive scan of the bin counts
			cntD = sBase[tid];
			inclD = cntD;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL_MASK, inclD, d); if (lane >= d) inclD += o; }
			if (lane == 31) sWarpSum[warp] = inclD;
		}
		__syncthreads();
		if (tid < 256) {
#pragma unroll
			for (uint32_t w = 0; w < 8; w++) if (w < warp) inclD += sWarpSum[w];
			sBase[tid] = inclD;
		}
		__syncthreads();
		totalAlive = sBase[255];
		if (tid < 256 && cntD) {
			// cut q sits behind the bin in which the running count reaches q / numParts of the keys
			for (uint32_t q = 1; q < numParts; q++) {
				const uint32_t target = (uint32_t)(((uint64_t)totalAlive * q + numParts - 1) / numParts);
				if (inclD >= target && inclD - cntD < target) sPivot[q] = tid + 1;
			}
		}
		__syncthreads();
		const uint32_t binLo = part ? sPivot[part] : 0u, binHi = part + 1 < numParts ? sPivot[part + 1] : 256u;
		if (binHi <= binLo) { if (part == 0 && tid == 0) S.aliveCount = totalAlive; return; } // an empty part (CTA-uniform)
		outOffset = binLo ? sBase[binLo - 1] : 0u;
		const uint32_t nMine = sBase[binHi - 1] - outOffset;
		if (nMine == 0) { if (part == 0 && tid == 0) S.aliveCount = totalAlive; return; }
		const uint32_t partBase = binLo << hshift;
		uint32_t mine = 0, mineBits = 0;
#pragma unroll
		for (uint32_t i = 0; i < kSmallItems; i++) {
			if (i >= items) break;
			if ((okBits >> i) & 1u) {
				const uint32_t b = key[i] >> hshift;
				if (b >= binLo && b < binHi) { mine++; mineBits |= 1u << i; }
			}
		}
		mine = __reduce_add_sync(FULL_MASK, mine);
		if (lane == 0) sPartWarp[warp] = mine;
		__syncthreads();
		uint32_t warpBase = 0;
#pragma unroll 8
		for (uint32_t w = 0; w < kSmallWarps; w++) if (w < warp) warpBase += sPartWarp[w];
		// ---- compact the part in ascending slot order = (warp, item, lane) order; its keys become relative to the part's base
		uint32_t run = warpBase;
#pragma unroll
		for (uint32_t i = 0; i < kSmallItems; i++) {
			if (i >= items) break;
			const bool in = (mineBits >> i) & 1u;
			const uint32_t bal = __ballot_sync(FULL_MASK, in);
			if (in) {
				const uint32_t pos = run + __popc(bal & ltMaskS);
				sKey[pos] = key[i] - partBase;
				sVal[pos] = (uint16_t)((i & 1u) ? (val2[i >> 1] >> 16) : (val2[i >> 1] & 0xffffu));
			}
			run += __popc(bal);
		}
		__syncthreads();
		items = (nMine + kSmallThreads - 1) / kSmallThreads;
		warpBegin = warp * 32 * items;
		okBits = 0;
#pragma unroll
		for (uint32_t i = 0; i < kSmallItems; i++) {
			if (i >= items) break;
			const uint32_t idx = warpBegin + i * 32 + lane;
			const bool ok = idx < nMine;
			key[i] = ok ? sKey[idx] : 0u;
			const uint32_t v = ok ? sVal[idx] : 0u;
			if (i & 1u) val2[i >> 1] |= v << 16; else val2[i >> 1] = v;
			if (ok) okBits |= 1u << i;
		}
		__syncthreads(); // all reads done before the first pass clears the tables aliased onto sKey
		const uint32_t span = ((binHi - binLo) << hshift) - 1u; // keys of the part, relative to its base, are <= span
		bits = span ? 32u - (uint32_t)__clz(span) : 0u;
	}
	const uint32_t numPasses = bits ? (bits + 7u) / 8u : 1u; // pass 0 always runs: it compacts

	uint32_t n = 0; // live keys; known after pass 0
	for (uint32_t pass = 0; pass < numPasses; pass++) {
		const uint32_t shift = pass * 8;
		// ---- clear the per-warp digit counts and the peer-mask tables (aliased onto sKey, idle while the keys are in registers)
		for (uint32_t i = tid; i < kSmallWarps * 256; i += kSmallThreads) { sCnt[i] = 0; sKey[i] = 0; sKey[i + kSmallWarps * 256] = 0; }
		__syncthreads();
		// ---- rank inside the warp (same scheme as k_sort_pass: one shared-memory atomicOr per lane finds the peers)
		uint32_t *cnt = sCnt + warp * 256;
		uint32_t *match0 = sKey + warp * 512, *match1 = match0 + 256;
		const uint32_t ltMask = (1u << lane) - 1u;
#pragma unroll
		for (uint32_t i = 0; i < kSmallItems; i++) {
			if (i >= items) break;
			uint32_t *match = (i & 1u) ? match1 : match0;
			const bool ok = (okBits >> i) & 1u;
			const uint32_t d = (key[i] >> shift) & 255u;
			if (ok) atomicOr(&match[d], 1u << lane);
			__syncwarp();
			const uint32_t peers = ok ? match[d] : 0u;
			const uint32_t before = ok ? cnt[d] : 0u;
			__syncwarp();
			const uint32_t rk = before + __popc(peers & ltMask);
			if (ok && (peers & ltMask) == 0) { cnt[d] = before + __popc(peers); match[d] = 0; }
			if (i & 1u) rank2[i >> 1] |= rk << 16; else rank2[i >> 1] = rk;
		}
		__syncthreads();
		// ---- thread d < 256: exclusive offsets of digit d over the warps, then the exclusive scan over the digits
		uint32_t total = 0;
		if (tid < 256) {
#pragma unroll 8
			for (uint32_t w = 0; w < kSmallWarps; w++) { const uint32_t c = sCnt[w * 256 + tid]; sCnt[w * 256 + tid] = total; total += c; }
			uint32_t incl = total;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL_MASK, incl, d); if (lane >= d) incl += o; }
			if (lane == 31) sWarpSum[warp] = incl;
			total = incl - total; // exclusive inside the warp
		}
		__syncthreads();
		if (tid < 256) {
			uint32_t warpBase = 0;
#pragma unroll
			for (uint32_t w = 0; w < 8; w++) if (w < warp) warpBase += sWarpSum[w];
			sBase[tid] = warpBase + total;
			if (tid == 255) { uint32_t all = 0; for (uint32_t w = 0; w < 8; w++) all += sWarpSum[w]; sTotal = all; }
		}
		__syncthreads();
		n = sTotal;
		// ---- permute through shared memory: position = digit base + this warp's offset + rank in the warp
		// (the peer-mask tables aliased onto sKey are dead since the barrier after the ranking loop)
#pragma unroll
		for (uint32_t i = 0; i < kSmallItems; i++) {
			if (i >= items) break;
			if ((okBits >> i) & 1u) {
				const uint32_t d = (key[i] >> shift) & 255u;
				const uint32_t rk = (i & 1u) ? (rank2[i >> 1] >> 16) : (rank2[i >> 1] & 0xffffu);
				const uint32_t pos = sBase[d] + sCnt[warp * 256 + d] + rk;
				if (pass + 1 < numPasses) sKey[pos] = key[i]; // after the last pass only the slot order is read
				sVal[pos] = (uint16_t)((i & 1u) ? (val2[i >> 1] >> 16) : (val2[i >> 1] & 0xffffu));
			}
		}
		__syncthreads();
		if (pass + 1 == numPasses) break;
		// ---- back into registers in the new order; after pass 0 the live keys are the first n
		okBits = 0;
#pragma unroll
		for (uint32_t i = 0; i < kSmallItems; i++) {
			if (i >= items) break;
			const uint32_t idx = warpBegin + i * 32 + lane;
			const bool ok = idx < n;
			key[i] = ok ? sKey[idx] : 0u;
			const uint32_t v = ok ? sVal[idx] : 0u;
			if (i & 1u) val2[i >> 1] |= v << 16; else val2[i >> 1] = v;
			if (ok) okBits |= 1u << i;
		}
		__syncthreads(); // all reads done before the next pass clears the aliased tables
	}
	// ---- the sorted slot list (PoolPtrs::vals: what k_expand_sorted and sprReadStreamSlots read)
	uint32_t *vout = pool.vals + segBase + outOffset;
	for (uint32_t j = tid; j < n; j += kSmallThreads) vout[j] = sVal[j];
	if (tid == 0 && part == 0) S.aliveCount = splitLog ? totalAlive : n; // gpuParticles.size / 4 (:670); the tile path writes it in k_sort_scan
}

// One lane per particle: a coalesced read of 32 slot indices, one 256-bit gather of the 32-byte
// record per lane, then the record is stored four times (:607-665). The four 256-bit stores of a
// warp together fill 32 complete 128-byte lines (transposed across the warp, see below).
// A plain load that misses makes L2 fetch the whole 128-byte line from DRAM (4 sectors for one 32-byte
// record: ncu lts__t_sectors_srcunit_tex_op_read = 4 per record, dram__bytes_read = 121 B per record).
// That is what a small effect wants -- its few MB of state stay in L2 while it is gathered, so the other three
// records of the line are hits a moment later -- but for an effect much larger than L2 they are evicted before
// they are asked for. There the L2::64B prefetch-size hint (the smallest PTX offers) halves the DRAM read:
// 1.21 -> 0.66 GB for 10M records (tools/gather_bench.cu, profiles/r01_gather_microbench.txt).
__device__ inline void ld_record_cg(const float4 *p, float4 &a, float4 &b)
{
	asm("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ inline void ld_record_64b(const float4 *p, float4 &a, float4 &b)
{
	asm("ld.global.L2::64B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
static const uint32_t kLargeGatherSlots = 2u << 20; // 64 MB of slot state: beyond this the gather asks for half lines
__device__ inline void st_record_cs(float4 *p, const float4 &a, const float4 &b)
{
	asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
		:: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// EXPAND = false (deferred expansion, sprContextSetDeferredExpansion): only the sorted slots are brought home to
// PoolPtrs::vals; the records are gathered later by the pack kernels, straight into a run buffer.
// One record per thread, 1024 threads per CTA, two CTAs per SM (32 registers, every warp slot of the SM in use) measured
// 1 % faster on the kernel than four records per thread in 256-thread CTAs (64 registers, half the warp slots) at C3, C4 and C2;
// two per thread at three CTAs of 512 (40 registers, spills) 4 % slower at C3: profiles/r02_sort_experiments.md M.
#ifndef SPR_EXPAND_PER_THREAD
#define SPR_EXPAND_PER_THREAD 1
#endif
#ifndef SPR_EXPAND_MIN_CTAS
#define SPR_EXPAND_MIN_CTAS 2
#endif
static const uint32_t kExpPer = SPR_EXPAND_PER_THREAD;       // records per thread of k_expand_sorted (all gathers in flight at once)
static const uint32_t kExpThreads = 1024 / kExpPer;          // a CTA expands 1024 consecutive ranks
template <bool EXPAND>
__global__ void __launch_bounds__(kExpThreads, SPR_EXPAND_MIN_CTAS) k_expand_sorted(PoolPtrs pool, TablePtrs tab, const uint32_t *segEffects, const SortTile *tiles,
	const uint32_t *valsAlt, const uint32_t *passMask)
{
	pdl_enter();
	// kExpandPerTile CTAs per sort tile: 1024 particles per CTA, kExpPer per thread, all gathers in flight at once
	SortTile st = tiles[blockIdx.x / kExpandPerTile];
	uint32_t g = segEffects[st.effect];
	uint32_t alive = tab.state[g].aliveCount;
	uint32_t begin = st.tile * kSortTileKeys + (blockIdx.x % kExpandPerTile) * 1024u;
	if (begin >= alive) return;
	uint32_t count = min(1024u, alive - begin);
	uint32_t segBase = tab.params[g].slotBase;
	// an odd number of executed passes leaves the sorted slots in the Alt buffer: copy them to PoolPtrs::vals on
	// the way (the parity read-back sprReadStreamSlots reads them there)
	const bool fromAlt = __popc(passMask[st.effect]) & 1u;
	if (!EXPAND && !fromAlt) return;
	const uint32_t *vals = (fromAlt ? valsAlt : pool.vals) + segBase + begin;
	const float4 *state = pool.state + 2ull * segBase;
	float4 *dst = pool.vertices + 8ull * (segBase + begin);
	uint32_t slot[kExpPer];
	float4 a[kExpPer], b[kExpPer];
#pragma unroll
	for (int u = 0; u < (int)kExpPer; u++) {
		a[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); b[u] = a[u];
		uint32_t i = u * kExpThreads + threadIdx.x;
		slot[u] = i < count ? vals[i] : 0xffffffffu;
		if (fromAlt && i < count) pool.vals[segBase + begin + i] = slot[u];
	}
	if (!EXPAND) return;
	if (tab.state[g].slotCount[tab.state[g].parity] > kLargeGatherSlots) {
#pragma unroll
		for (int u = 0; u < (int)kExpPer; u++) if (slot[u] != 0xffffffffu) ld_record_64b(state + 2ull * slot[u], a[u], b[u]);
	} else {
#pragma unroll
		for (int u = 0; u < (int)kExpPer; u++) if (slot[u] != 0xffffffffu) ld_record_cg(state + 2ull * slot[u], a[u], b[u]);
	}
	// The warp's 32 records go out transposed: store j of lane l is copy (l & 3) of the group's particle j * 8 + l / 4,
	// fetched from that particle's lane with shuffles, so that each store instruction of the warp covers 1 KB of
	// consecutive bytes (8 lines) instead of one 32-byte piece of 32 different lines. Same bytes, a quarter of the
	// LSU wavefronts: tools/expand_store_bench.cu, 10M random 0.392 -> 0.382 ms, 1M (L2 resident) 0.042 -> 0.034 ms.
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
	for (int u = 0; u < (int)kExpPer; u++) {
		const uint32_t wbase = u * kExpThreads + warp * 32;   // the warp's group of 32 consecutive ranks
		if (wbase >= count) continue;                   // warp-uniform
		const uint32_t cnt = min(32u, count - wbase);
		float4 *d = dst + 8ull * wbase;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const uint32_t p = j * 8 + (lane >> 2);
			float4 x, y;
			x.x = __shfl_sync(FULL_MASK, a[u].x, p); x.y = __shfl_sync(FULL_MASK, a[u].y, p); x.z = __shfl_sync(FULL_MASK, a[u].z, p); x.w = __shfl_sync(FULL_MASK, a[u].w, p);
			y.x = __shfl_sync(FULL_MASK, b[u].x, p); y.y = __shfl_sync(FULL_MASK, b[u].y, p); y.z = __shfl_sync(FULL_MASK, b[u].z, p); y.w = __shfl_sync(FULL_MASK, b[u].w, p);
			if (p < cnt) st_record_cs(d + 2 * (j * 32 + lane), x, y);
		}
	}
}

// Per-device kernel attributes; called once per context: 4 CTAs x 43 KB of shared memory per SM.
void configure_sort_kernels()
{
	cudaFuncSetAttribute(k_sort_pass<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
	cudaFuncSetAttribute(k_sort_pass<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
	cudaFuncSetAttribute(k_sort_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPassSmemBytes);
	cudaFuncSetAttribute(k_sort_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPassSmemBytes);
	cudaFuncSetAttribute(k_sort_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSmemBytes);
}

void clear_sort_scratch(cudaStream_t st, const SortScratch &sc, uint32_t numSegments, uint32_t numTiles)
{
	if (!numSegments || !numTiles) return;
	cudaMemsetAsync(sc.hist, 0, ((size_t)numSegments * 1025 + 4) * sizeof(uint32_t), st); // histograms, pass masks, the four pass tickets
	cudaMemsetAsync(sc.lookback, 0, (size_t)numTiles * 4 * 256 * sizeof(uint32_t), st);
}

uint32_t launch_sort_and_expand(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, const SortScratch &sc,
	const uint32_t *segEffects, uint32_t numSegments, uint32_t numBigSegments, const SortTile *tiles, uint32_t numTiles, uint32_t numBigTiles,
	const SortTimingHooks *hooks)
{
	if (!numSegments || !numTiles) return 0;
	void *tok = nullptr;
#define T_BEGIN(k) do { if (hooks) tok = hooks->begin(hooks->ctx, k); } while (0)
#define T_END(k) do { if (hooks) hooks->end(hooks->ctx, k, tok); } while (0)
	uint32_t *passMask = sc.hist + (size_t)numSegments * 1024; // [numSegments], cleared together with the histograms (clear_sort_scratch)
	uint32_t *tickets = passMask + numSegments;                 // [4] one tile ticket per pass, cleared with them
	uint32_t launches = 0;
	if (numBigSegments && numBigTiles) {
		T_BEGIN(0);
		launch_chain(k_sort_hist, numBigTiles, 256, 0, st, pool, tab, segEffects, tiles, sc.hist);
		T_END(0);
		uint32_t rows = numBigSegments * 4;
		T_BEGIN(1);
		launch_chain(k_sort_scan, (rows * 32 + 255) / 256, 256, 0, st, tab, segEffects, sc.hist, passMask, rows);
		T_END(1);
		uint32_t *kbuf[2] = { pool.keys, sc.keysAlt };
		uint32_t *vbuf[2] = { pool.vals, sc.valsAlt };
		for (uint32_t p = 0; p < 4; p++) {
			T_BEGIN(2);
			uint32_t *lb = sc.lookback + (size_t)p * numBigTiles * 256;
			if (p == 0) launch_chain(k_sort_pass<true>, numBigTiles, kSortThreads, kPassSmemBytes, st, pool, tab, segEffects, tiles, kbuf[0], vbuf[0], kbuf[1], vbuf[1], sc.hist, passMask, lb, tickets + p, p);
			else launch_chain(k_sort_pass<false>, numBigTiles, kSortThreads, kPassSmemBytes, st, pool, tab, segEffects, tiles, kbuf[0], vbuf[0], kbuf[1], vbuf[1], sc.hist, passMask, lb, tickets + p, p);
			T_END(2);
		}
		launches += 6;
	}
	if (numSegments > numBigSegments) {
		// their pass masks stay 0 (cleared with the histograms): k_expand_sorted then reads PoolPtrs::vals
		T_BEGIN(4);
		// few one-CTA effects, at least one of them big enough to be worth it: 2 or 4 CTAs share each effect by key range
		static const bool splitAllowed = [] { const char *e = getenv("SPR_SMALL_SPLIT"); return !(e && e[0] == '0'); }(); // A/B: SPR_SMALL_SPLIT=0
		const uint32_t numSmall = numSegments - numBigSegments;
		uint32_t splitLog = 0;
		if (splitAllowed && sc.smallMaxSlots > 4096u) splitLog = numSmall * 4 <= sc.numSms ? 2u : numSmall * 2 <= sc.numSms ? 1u : 0u;
		launch_chain(k_sort_small, numSmall << splitLog, kSmallThreads, kSmallSmemBytes, st, pool, tab, segEffects + numBigSegments, splitLog);
		T_END(4);
		launches++;
	}
	T_BEGIN(3);
	if (sc.deferExpansion) launch_chain(k_expand_sorted<false>, numTiles * kExpandPerTile, kExpThreads, 0, st, pool, tab, segEffects, tiles, sc.valsAlt, passMask);
	else launch_chain(k_expand_sorted<true>, numTiles * kExpandPerTile, kExpThreads, 0, st, pool, tab, segEffects, tiles, sc.valsAlt, passMask);
	T_END(3);
	return launches + 1;
}

} // namespace spr
