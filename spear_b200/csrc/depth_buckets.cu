// depth_buckets.cu -- the sorted step of a large effect without a random gather over its whole state.
//
// Extension row A11/K4 of SURVEY.md (per-particle depth order; idiom src/client/BillboardSystem.cpp:41-51: depth
// descending, ties by ascending slot), same result as the radix path of sort.cu, bit for bit.
//
// sort.cu sorts (key, slot) pairs and then gathers the 32-byte records in depth order: 10M random 32-byte reads out
// of 320 MB, bound by the random-access rate of the memory system (profiles/r01_gather_microbench.txt). Particles move
// little per frame, so the depth QUANTILES of the last frame ("splitters", one per ~4600 particles) still cut this
// frame's particles into buckets of about that size. The records themselves are partitioned by bucket while they are
// read sequentially (a stable counting partition: slot order survives inside a bucket), and one CTA finishes a bucket:
// recompute the keys from the records, stable LSD sort of (key, index in the bucket) in shared memory, 4x expansion
// (:607-665) with the gather confined to the bucket's 150 KB, which it has just read -- L2 hits. Any monotone function
// of the key is a correct bucket function, so stale splitters only unbalance the buckets; a bucket above the CTA's
// capacity makes the whole effect take the radix path for that frame (k_bucket_scan_bins decides, on the device).
//
//   k_bucket_hist        per 4096-slot tile: key -> bucket (binary search of the splitters), u16 bucket id per slot,
//                        per-tile bucket counts
//   k_bucket_scan_tiles  per chunk of 64 tiles x 32 buckets: exclusive scan over the tiles, chunk totals
//   k_bucket_scan_bins   per effect: scan over the chunks and over the buckets -> bucket bases and sizes, live count,
//                        verdict (every bucket fits), next frame's splitter table pre-filled
//   k_bucket_scatter     per tile: stable rank of every live slot inside (tile, bucket) with the peer-mask scheme of
//                        k_sort_pass, record moved to bucket base + chunk offset + tile offset + rank
//   k_bucket_finish      per bucket: keys from the records, LSD passes in shared memory, expansion, next splitters
//   k_bucket_splitters   per effect the radix path sorted: next frame's splitters from the sorted slots

#include <cuda_runtime.h>
#include <stdint.h>

#include "device_types.h"
#include "kernels.h"
#include "launch.h"

namespace spr {

#define FULL_MASK 0xffffffffu

static const uint32_t kTileKeys = kSortTileKeys;       // slots per tile (the tile list is shared with sort.cu)
static const uint32_t kTileItems = kTileKeys / 256;    // slots per thread of a 256-thread tile CTA
static const uint32_t kTileWarpKeys = 32 * kTileItems; // slots per warp
static_assert(kTileKeys == 4096, "the depth-bucket kernels are written for 4096-slot tiles (16-bit tile-local offsets, 8 warps x 512)");

__device__ inline uint32_t pow2ceil_dev(uint32_t v) { return v <= 1 ? 1u : 1u << (32 - __clz(v - 1)); }

__device__ inline void bk_ld_record(const float4 *p, float4 &a, float4 &b)
{
	asm("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ inline void bk_ld_record_cg(const float4 *p, float4 &a, float4 &b)
{
	asm("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
// L2 policies for k_bucket_finish: a bucket's records are read twice by its CTA (keys, then the gather in depth order)
// with the bucket's share of 1.3 GB of expansion stores in between -- the records stay (evict_last), the stores go first.
__device__ inline unsigned long long l2_policy_evict_last()
{
	unsigned long long pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
__device__ inline unsigned long long l2_policy_evict_first()
{
	unsigned long long pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
__device__ inline float4 bk_ld_f4_keep(const float4 *p, unsigned long long pol)
{
	float4 v;
	asm("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
	return v;
}
__device__ inline void bk_st_record_pol(float4 *p, const float4 &a, const float4 &b, unsigned long long pol)
{
	asm volatile("st.global.L2::cache_hint.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8}, %9;"
		:: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w), "l"(pol) : "memory");
}
__device__ inline void bk_st_record(float4 *p, const float4 &a, const float4 &b)
{
	asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
		:: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}
__device__ inline void bk_st_record_cs(float4 *p, const float4 &a, const float4 &b)
{
	asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
		:: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// k_bucket_hist
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bucket_hist(PoolPtrs pool, TablePtrs tab, const uint32_t *segEffects, const SortTile *tiles,
	const BucketSeg *segs, const uint32_t *splitters, const uint16_t *lut, const uint32_t *lutMeta, uint16_t *binId, uint16_t *tileCnt, uint32_t stride)
{
	pdl_enter();
	__shared__ uint32_t sSpl[kMaxBins];
	__shared__ uint32_t sHist[kMaxBins];
	__shared__ uint16_t sLut[kLutCells + 2];
	const SortTile st = tiles[blockIdx.x];
	const BucketSeg seg = segs[st.effect];
	if (!seg.bins) return;
	const uint32_t g = segEffects[st.effect];
	const EffectState &S = tab.state[g];
	const uint32_t slots = S.slotCount[S.parity];
	const uint32_t begin = st.tile * kTileKeys;
	uint16_t *row = tileCnt + (size_t)blockIdx.x * stride;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	if (begin >= slots) { // a tile behind the effect's last slot (the list covers its slot BOUND): an empty row
		for (uint32_t b = tid; b < stride; b += 256) row[b] = 0;
		return;
	}
	const uint32_t count = min(kTileKeys, slots - begin);
	const uint32_t segBase = tab.params[g].slotBase;
	const uint32_t *keys = pool.keys + segBase + begin;
	const uint32_t *mask = pool.aliveMask + ((segBase + begin) >> 5);
	// all loads of the tile first, no branch between them (as k_sort_hist): key i = tid + 256 * it sits in alive word
	// warp + 8 * it, bit lane; a dead slot's key is stale but mapped
	uint32_t myWord = 0;
	if (lane < kTileItems && (warp + 8 * lane) * 32 < count) myWord = mask[warp + 8 * lane];
	uint32_t key[kTileItems];
#pragma unroll
	for (uint32_t it = 0; it < kTileItems; it++) {
		const uint32_t i = tid + it * 256u;
		key[it] = i < count ? keys[i] : 0u;
	}
	// the splitter table (entry 0 is 0, entry b the smallest key of bucket b) and its look-up table: cell c of the key
	// range [lo, hi) holds the bucket of the cell's first key, so a key's bucket lies in [lut[c], lut[c + 1]]
	const uint32_t *spl = splitters + ((size_t)seg.table * 2 + seg.parity) * kMaxBins;
	const uint16_t *lutG = lut + ((size_t)seg.table * 2 + seg.parity) * (kLutCells + 2);
	const uint32_t *meta = lutMeta + ((size_t)seg.table * 2 + seg.parity) * 4;
	for (uint32_t b = tid; b < seg.bins; b += 256) sSpl[b] = spl[b];
	for (uint32_t c = tid; c < kLutCells + 1; c += 256) sLut[c] = lutG[c];
	for (uint32_t b = tid; b < stride; b += 256) sHist[b] = 0;
	const uint32_t lo = meta[0], shift = meta[1];
	__syncthreads();
	uint16_t *ids = binId + segBase + begin;
#pragma unroll
	for (uint32_t it = 0; it < kTileItems; it++) {
		const uint32_t i = tid + it * 256u;
		const uint32_t word = __shfl_sync(FULL_MASK, myWord, it);
		const bool ok = i < count && ((word >> lane) & 1u);
		// bucket = (number of splitters <= key) - 1
		uint32_t pos = 0;
		if (key[it] >= lo) {
			const uint32_t c = min((key[it] - lo) >> shift, kLutCells - 1u);
			uint32_t a = sLut[c], b = sLut[c + 1];
			while (a < b) { const uint32_t mid = (a + b + 1) >> 1; if (sSpl[mid] <= key[it]) a = mid; else b = mid - 1; }
			pos = a;
		}
		if (ok) atomicAdd(&sHist[pos], 1u);
		if (i < count) ids[i] = ok ? (uint16_t)pos : (uint16_t)0xffffu;
	}
	__syncthreads();
	for (uint32_t b = tid; b < stride; b += 256) row[b] = (uint16_t)sHist[b];
}

// ---------------------------------------------------------------------------------------------------------------
// k_bucket_scan_tiles: one warp per (chunk, group of 32 buckets), lanes = buckets
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bucket_scan_tiles(const BucketSeg *segs, const BucketChunk *chunks, uint32_t numChunks,
	const uint16_t *tileCnt, uint32_t *tileOff, uint32_t *chunkSum, uint32_t stride)
{
	pdl_enter();
	const uint32_t groups = stride / 32;
	const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
	if (w >= numChunks * groups) return;
	const BucketChunk ch = chunks[w / groups];
	if (!segs[ch.seg].bins) return;
	const uint32_t bin = (w % groups) * 32 + lane;
	const uint16_t *src = tileCnt + (size_t)ch.firstRow * stride + bin;
	uint32_t *dst = tileOff + (size_t)ch.firstRow * stride + bin;
	uint32_t run = 0;
	for (uint32_t t0 = 0; t0 < ch.numTiles; t0 += 8) {
		uint32_t c[8];
#pragma unroll
		for (uint32_t k = 0; k < 8; k++) c[k] = t0 + k < ch.numTiles ? src[(size_t)(t0 + k) * stride] : 0u;
#pragma unroll
		for (uint32_t k = 0; k < 8; k++) {
			if (t0 + k < ch.numTiles) dst[(size_t)(t0 + k) * stride] = run;
			run += c[k];
		}
	}
	chunkSum[(size_t)ch.row * stride + bin] = run;
}

// ---------------------------------------------------------------------------------------------------------------
// k_bucket_scan_bins: one 1024-thread CTA per segment. chunkSum becomes the exclusive offset of the chunk inside its
// bucket; bucketBase / bucketCount the bucket's place in the effect's partitioned records and in its sorted stream.
// ---------------------------------------------------------------------------------------------------------------
__device__ inline uint32_t block_excl_scan_1024(uint32_t v, uint32_t *sWarp, uint32_t &blockTotal)
{
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint32_t incl = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL_MASK, incl, d); if (lane >= d) incl += o; }
	__syncthreads(); // (sWarp may still be read from a previous call)
	if (lane == 31) sWarp[warp] = incl;
	__syncthreads();
	uint32_t base = 0, total = 0;
#pragma unroll 8
	for (uint32_t w = 0; w < 32; w++) { const uint32_t s = sWarp[w]; if (w < warp) base += s; total += s; }
	blockTotal = total;
	return base + incl - v;
}

__global__ void __launch_bounds__(1024) k_bucket_scan_bins(TablePtrs tab, const uint32_t *segEffects, const BucketSeg *segs,
	uint32_t *chunkSum, uint32_t *bucketBase, uint32_t *bucketCount, uint32_t *bucketOk, unsigned long long *stats, uint32_t *splitters, uint32_t stride)
{
	pdl_enter();
	__shared__ uint32_t sWarp[32];
	__shared__ uint32_t sMax;
	const uint32_t s = blockIdx.x, tid = threadIdx.x;
	const BucketSeg seg = segs[s];
	if (!seg.bins) return; // no splitters yet: bucketOk stays 0 and the radix path sorts the effect
	const uint32_t numChunks = (seg.numTiles + kChunkTiles - 1) / kChunkTiles;
	if (tid == 0) sMax = 0;
	uint32_t total[2] = { 0, 0 };
#pragma unroll
	for (uint32_t k = 0; k < 2; k++) {
		const uint32_t bin = tid + k * 1024u;
		if (bin >= stride) continue;
		uint32_t *col = chunkSum + (size_t)seg.firstChunk * stride + bin;
		uint32_t run = 0;
		for (uint32_t c0 = 0; c0 < numChunks; c0 += 8) {
			uint32_t v[8];
#pragma unroll
			for (uint32_t i = 0; i < 8; i++) v[i] = c0 + i < numChunks ? col[(size_t)(c0 + i) * stride] : 0u;
#pragma unroll
			for (uint32_t i = 0; i < 8; i++) { if (c0 + i < numChunks) col[(size_t)(c0 + i) * stride] = run; run += v[i]; }
		}
		total[k] = run;
	}
	uint32_t sum0 = 0, sum1 = 0;
	const uint32_t excl0 = block_excl_scan_1024(total[0], sWarp, sum0);
	const uint32_t excl1 = sum0 + block_excl_scan_1024(total[1], sWarp, sum1);
	const uint32_t big = max(total[0], total[1]);
	const uint32_t warpBig = __reduce_max_sync(FULL_MASK, big);
	if ((tid & 31u) == 0 && warpBig) atomicMax(&sMax, warpBig);
	if (tid < stride) { bucketBase[(size_t)s * stride + tid] = excl0; bucketCount[(size_t)s * stride + tid] = total[0]; }
	if (tid + 1024u < stride) { bucketBase[(size_t)s * stride + tid + 1024u] = excl1; bucketCount[(size_t)s * stride + tid + 1024u] = total[1]; }
	__syncthreads();
	if (tid == 0) {
		tab.state[segEffects[s]].aliveCount = sum0 + sum1; // gpuParticles.size / 4 (:670)
		const bool fits = sMax <= kBucketCap;
		bucketOk[s] = fits ? 1u : 0u;
		atomicAdd(&stats[fits ? 0 : 1], 1ull);
	}
	// next frame's table: entry 0 and the padding here, the quantiles by k_bucket_finish (or, if the verdict is no,
	// all of it by k_bucket_splitters)
	if (seg.nextBins) {
		uint32_t *next = splitters + ((size_t)seg.table * 2 + (seg.parity ^ 1u)) * kMaxBins;
		const uint32_t padded = pow2ceil_dev(seg.nextBins);
		for (uint32_t b = tid; b < padded; b += 1024) next[b] = b ? 0xffffffffu : 0u;
	}
}

// ---------------------------------------------------------------------------------------------------------------
// k_bucket_scatter: the stable partition of one tile's live records. 128 threads: warp w owns slots [w * 1024, +1024)
// of the tile, 32 per lane -- four per-warp tables instead of eight keep the CTA at 28 bytes of shared memory per bucket
// (4 CTAs per SM at 2048 buckets).
// dynamic shared memory: sOff u32[stride] | sMatch u32[4][stride] | sCnt u16[4][stride]
// ---------------------------------------------------------------------------------------------------------------
static const uint32_t kScThreads = 128;
static const uint32_t kScWarps = kScThreads / 32;
static const uint32_t kScItems = kTileKeys / kScThreads;   // 32 slots per thread
static const uint32_t kScWarpKeys = 32 * kScItems;         // 1024 slots per warp

__global__ void __launch_bounds__(kScThreads, 4) k_bucket_scatter(PoolPtrs pool, TablePtrs tab, const uint32_t *segEffects, const SortTile *tiles,
	const BucketSeg *segs, const uint32_t *bucketOk, const uint16_t *binId, const uint32_t *tileOff, const uint32_t *chunkOff,
	const uint32_t *bucketBase, float4 *part, uint32_t stride, uint32_t dbg)
{
	pdl_enter();
	extern __shared__ __align__(16) uint32_t sDyn[];
	uint32_t *sOff = sDyn;
	uint32_t *sMatch = sDyn + stride;
	uint16_t *sCnt = (uint16_t*)(sDyn + (1 + kScWarps) * (size_t)stride);
	const SortTile st = tiles[blockIdx.x];
	const BucketSeg seg = segs[st.effect];
	if (!seg.bins || !bucketOk[st.effect]) return;
	const uint32_t g = segEffects[st.effect];
	const EffectState &S = tab.state[g];
	const uint32_t slots = S.slotCount[S.parity];
	const uint32_t begin = st.tile * kTileKeys;
	if (begin >= slots) return;
	const uint32_t count = min(kTileKeys, slots - begin);
	const uint32_t segBase = tab.params[g].slotBase;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

	// bucket ids of the warp's 1024 slots (item i of lane l is slot warp * 1024 + i * 32 + l), two per register
	const uint16_t *ids = binId + segBase + begin;
	uint32_t id2[kScItems / 2];
#pragma unroll
	for (uint32_t i = 0; i < kScItems; i++) {
		const uint32_t idx = warp * kScWarpKeys + i * 32 + lane;
		const uint32_t v = idx < count ? (uint32_t)ids[idx] : 0xffffu;
		if (i & 1u) id2[i >> 1] |= v << 16; else id2[i >> 1] = v;
	}
	// where this tile's records of bucket b start: bucket base + chunk offset + tile offset
	const uint32_t chunkRow = seg.firstChunk + st.tile / kChunkTiles;
	{
		const uint32_t *bb = bucketBase + (size_t)st.effect * stride;
		const uint32_t *co = chunkOff + (size_t)chunkRow * stride;
		const uint32_t *to = tileOff + (size_t)blockIdx.x * stride;
		for (uint32_t b0 = tid; b0 < seg.bins; b0 += 4 * kScThreads) {
			uint32_t v[4];
#pragma unroll
			for (uint32_t u = 0; u < 4; u++) { const uint32_t b = b0 + u * kScThreads; v[u] = b < seg.bins ? bb[b] + co[b] + to[b] : 0u; }
#pragma unroll
			for (uint32_t u = 0; u < 4; u++) { const uint32_t b = b0 + u * kScThreads; if (b < seg.bins) sOff[b] = v[u]; }
		}
	}
	{
		// clear the peer masks and the per-warp counts (contiguous behind sOff)
		uint4 *z = (uint4*)(sDyn + stride);
		const uint32_t n16 = stride * (kScWarps * 6u) / 16u;
		for (uint32_t i = tid; i < n16; i += kScThreads) z[i] = make_uint4(0, 0, 0, 0);
	}
	__syncthreads();

	// rank inside (warp, bucket): peers of a lane = lanes of the warp with the same bucket in this round, found with one
	// shared-memory atomicOr per lane (k_sort_pass); the round's leader adds the round to the warp's count of the bucket
	uint32_t *match = sMatch + warp * stride;
	uint16_t *cnt = sCnt + warp * stride;
	const uint32_t ltMask = (1u << lane) - 1u;
	uint32_t rank2[kScItems / 2];
#pragma unroll
	for (uint32_t i = 0; i < kScItems; i++) {
		const uint32_t idv = (i & 1u) ? (id2[i >> 1] >> 16) : (id2[i >> 1] & 0xffffu);
		const bool ok = idv != 0xffffu;
		const uint32_t d = ok ? idv : 0u;
		if (dbg & 2u) { if (i & 1u) rank2[i >> 1] |= 0u; else rank2[i >> 1] = 0u; continue; }
		if (ok) atomicOr(&match[d], 1u << lane);
		__syncwarp();
		const uint32_t peers = ok ? match[d] : 0u;
		const uint32_t before = ok ? (uint32_t)cnt[d] : 0u;
		__syncwarp();
		const uint32_t rk = before + __popc(peers & ltMask);
		if (ok && (peers & ltMask) == 0) { cnt[d] = (uint16_t)(before + __popc(peers)); match[d] = 0; }
		__syncwarp(); // the leader's clear lands before the next round's atomicOr on the same entry
		if (i & 1u) rank2[i >> 1] |= rk << 16; else rank2[i >> 1] = rk;
	}
	__syncthreads();
	// per bucket: the warps' counts become their exclusive offsets inside the tile
	for (uint32_t b = tid; b < seg.bins; b += kScThreads) {
		uint32_t run = 0;
#pragma unroll
		for (uint32_t w = 0; w < kScWarps; w++) { const uint32_t c = sCnt[w * stride + b]; sCnt[w * stride + b] = (uint16_t)run; run += c; }
	}
	__syncthreads();

	// move the records: coalesced 256-bit reads in slot order, one whole 32-byte sector written per record
	const float4 *src = pool.state + 2ull * (segBase + begin);
	float4 *dstBase = part + 2ull * segBase;
#pragma unroll
	for (uint32_t i0 = 0; i0 < kScItems; i0 += 4) {
		float4 a[4], b[4];
		uint32_t dst[4];
#pragma unroll
		for (uint32_t u = 0; u < 4; u++) {
			const uint32_t i = i0 + u;
			const uint32_t idv = (i & 1u) ? (id2[i >> 1] >> 16) : (id2[i >> 1] & 0xffffu);
			dst[u] = 0xffffffffu;
			if (idv != 0xffffu) {
				const uint32_t rk = (i & 1u) ? (rank2[i >> 1] >> 16) : (rank2[i >> 1] & 0xffffu);
				dst[u] = sOff[idv] + (uint32_t)cnt[idv] + rk;
				bk_ld_record(src + 2ull * (warp * kScWarpKeys + i * 32 + lane), a[u], b[u]);
			}
		}
#pragma unroll
		for (uint32_t u = 0; u < 4; u++) if (dst[u] != 0xffffffffu && !(dbg & 1u)) bk_st_record(dstBase + 2ull * dst[u], a[u], b[u]);
	}
}

// ---------------------------------------------------------------------------------------------------------------
// k_bucket_finish: one 512-thread CTA per bucket (at most kBucketCap records, 18 per thread).
// ---------------------------------------------------------------------------------------------------------------
static const uint32_t kFinThreads = 512;
static const uint32_t kFinWarps = kFinThreads / 32;
static const uint32_t kFinItems = kBucketCap / kFinThreads; // 18
// shared memory: keys (aliased by the per-warp peer-mask tables while ranking), 16-bit indices, per-warp digit counts
static const size_t kFinSmemBytes = kBucketCap * 4 + kBucketCap * 2 + kFinWarps * 256 * 4;
static_assert(kFinItems * kFinThreads == kBucketCap, "kBucketCap must be a multiple of the CTA size");
static_assert(kFinWarps * 512 * 4 <= kBucketCap * 4, "the peer-mask tables alias the key buffer");
static_assert(kBucketCap <= 65536, "indices inside a bucket are 16-bit");

__global__ void __launch_bounds__(kFinThreads, 2) k_bucket_finish(PoolPtrs pool, TablePtrs tab, const uint32_t *segEffects, const BucketSeg *segs,
	const BucketJob *jobs, const uint32_t *bucketOk, const uint32_t *bucketBase, const uint32_t *bucketCount, const float4 *part,
	uint32_t *splitters, uint32_t stride, float camX, float camY, float camZ, uint32_t dbg)
{
	pdl_enter();
	extern __shared__ __align__(16) uint8_t sRaw[];
	const unsigned long long polKeep = l2_policy_evict_last(), polStream = l2_policy_evict_first();
	uint32_t *sKey = (uint32_t*)sRaw;                               // [kBucketCap]
	uint16_t *sVal = (uint16_t*)(sRaw + kBucketCap * 4);            // [kBucketCap]
	uint32_t *sCnt = (uint32_t*)(sRaw + kBucketCap * 6);            // [kFinWarps][256]
	__shared__ uint32_t sBase[256];
	__shared__ uint32_t sWarpSum[8];
	__shared__ uint32_t sMinKey, sMaxKey;

	const BucketJob job = jobs[blockIdx.x];
	const BucketSeg seg = segs[job.seg];
	if (job.bin >= seg.bins || !bucketOk[job.seg]) return;
	const uint32_t n = bucketCount[(size_t)job.seg * stride + job.bin];
	if (!n) return;
	if (n > kBucketCap) __trap(); // k_bucket_scan_bins said every bucket fits
	const uint32_t base = bucketBase[(size_t)job.seg * stride + job.bin];
	const uint32_t g = segEffects[job.seg];
	const uint32_t segBase = tab.params[g].slotBase;
	const float4 *rec = part + 2ull * (segBase + base);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
	// warp-striped: warp w owns [w * 32 * items, (w + 1) * 32 * items), item i of lane l is w * 32 * items + i * 32 + l,
	// ascending in (warp, item, lane) = the order ranks are given in = arrival order = slot order
	const uint32_t items = (n + kFinThreads - 1) / kFinThreads;
	const uint32_t warpBegin = warp * 32 * items;

	if (tid == 0) { sMinKey = 0xffffffffu; sMaxKey = 0; }
	__syncthreads();

	// ---- keys from the records (the first 16 bytes of each; the whole sector comes into L2 for the gather below)
	uint32_t key[kFinItems], rank2[kFinItems / 2], val2[kFinItems / 2];
	uint32_t okBits = 0, mn = 0xffffffffu, mx = 0;
	// (six loads in flight per thread: the batch's loads are issued before anything depends on one)
#pragma unroll
	for (uint32_t i0 = 0; i0 < kFinItems; i0 += 6) {
		float4 p[6];
#pragma unroll
		for (uint32_t u = 0; u < 6; u++) {
			const uint32_t idx = warpBegin + (i0 + u) * 32 + lane;
			p[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			if (i0 + u < items && idx < n) p[u] = (dbg & 16u) ? __ldcg(rec + 2ull * idx) : bk_ld_f4_keep(rec + 2ull * idx, polKeep);
		}
#pragma unroll
		for (uint32_t u = 0; u < 6; u++) {
			const uint32_t i = i0 + u;
			const uint32_t idx = warpBegin + i * 32 + lane;
			key[i] = 0;
			if (i < items && idx < n) {
				key[i] = depth_key(p[u].x, p[u].y, p[u].z, camX, camY, camZ);
				okBits |= 1u << i;
				mn = min(mn, key[i]); mx = max(mx, key[i]);
			}
			const uint32_t v = idx & 0xffffu;
			if (i & 1u) val2[i >> 1] |= v << 16; else val2[i >> 1] = v;
		}
	}
	mn = __reduce_min_sync(FULL_MASK, mn); mx = __reduce_max_sync(FULL_MASK, mx);
	if (lane == 0) { atomicMin(&sMinKey, mn); atomicMax(&sMaxKey, mx); }
	__syncthreads();
	const uint32_t keyMin = sMinKey;
	const uint32_t maxRel = sMaxKey - keyMin;
	const uint32_t numPasses = (maxRel && !(dbg & 8u)) ? (32u - (uint32_t)__clz(maxRel) + 7u) / 8u : 0u;
#pragma unroll
	for (uint32_t i = 0; i < kFinItems; i++) { if (i >= items) break; key[i] -= keyMin; }
	if (numPasses == 0) {
		// every key is the same: the arrival order is the sorted order
		for (uint32_t j = tid; j < n; j += kFinThreads) { sVal[j] = (uint16_t)j; sKey[j] = 0; }
		__syncthreads();
	}

	for (uint32_t pass = 0; pass < numPasses; pass++) {
		const uint32_t shift = pass * 8;
		// ---- clear the per-warp digit counts and the peer-mask tables (aliased onto sKey, idle while the keys are in registers)
		for (uint32_t i = tid; i < kFinWarps * 256; i += kFinThreads) { sCnt[i] = 0; sKey[i] = 0; sKey[i + kFinWarps * 256] = 0; }
		__syncthreads();
		uint32_t *cnt = sCnt + warp * 256;
		uint32_t *match0 = sKey + warp * 512, *match1 = match0 + 256;
		const uint32_t ltMask = (1u << lane) - 1u;
#pragma unroll
		for (uint32_t i = 0; i < kFinItems; i++) {
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
			for (uint32_t w = 0; w < kFinWarps; w++) { const uint32_t c = sCnt[w * 256 + tid]; sCnt[w * 256 + tid] = total; total += c; }
			uint32_t incl = total;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL_MASK, incl, d); if (lane >= d) incl += o; }
			if (lane == 31) sWarpSum[warp] = incl;
			total = incl - total;
		}
		__syncthreads();
		if (tid < 256) {
			uint32_t warpBase = 0;
#pragma unroll
			for (uint32_t w = 0; w < 8; w++) if (w < warp) warpBase += sWarpSum[w];
			sBase[tid] = warpBase + total;
		}
		__syncthreads();
		// ---- permute through shared memory (the peer-mask tables aliased onto sKey are dead since the barrier after the ranking loop)
#pragma unroll
		for (uint32_t i = 0; i < kFinItems; i++) {
			if (i >= items) break;
			if ((okBits >> i) & 1u) {
				const uint32_t d = (key[i] >> shift) & 255u;
				const uint32_t rk = (i & 1u) ? (rank2[i >> 1] >> 16) : (rank2[i >> 1] & 0xffffu);
				const uint32_t pos = sBase[d] + sCnt[warp * 256 + d] + rk;
				sKey[pos] = key[i];
				sVal[pos] = (uint16_t)((i & 1u) ? (val2[i >> 1] >> 16) : (val2[i >> 1] & 0xffffu));
			}
		}
		__syncthreads();
		if (pass + 1 == numPasses) break;
		// ---- back into registers in the new order (every item of the bucket is live: the first n stay the first n)
#pragma unroll
		for (uint32_t i = 0; i < kFinItems; i++) {
			if (i >= items) break;
			const uint32_t idx = warpBegin + i * 32 + lane;
			const bool ok = idx < n;
			key[i] = ok ? sKey[idx] : 0u;
			const uint32_t v = ok ? sVal[idx] : 0u;
			if (i & 1u) val2[i >> 1] |= v << 16; else val2[i >> 1] = v;
		}
		__syncthreads(); // all reads done before the next pass clears the aliased tables
	}

	// ---- next frame's splitters: entry j is the key of stream rank floor(j * N / nextBins); this bucket holds ranks [base, base + n)
	if (seg.nextBins) {
		const unsigned long long N = tab.state[g].aliveCount, nb = seg.nextBins;
		uint32_t *next = splitters + ((size_t)seg.table * 2 + (seg.parity ^ 1u)) * kMaxBins;
		unsigned long long j = ((unsigned long long)base * nb + N - 1) / N; // smallest j with floor(j * N / nb) >= base
		if (j == 0) j = 1;
		for (j += tid; j < nb; j += kFinThreads) {
			const unsigned long long r = j * N / nb;
			if (r >= (unsigned long long)base + n) break;
			next[j] = keyMin + sKey[(uint32_t)r - base];
		}
	}

	// ---- expansion in depth order (:607-665): rank j of the bucket is stream rank base + j. The gather goes to the
	// bucket's own records, read a moment ago; the stores are transposed across the warp as in k_expand_sorted.
	// The four stores of a warp's group of 32 ranks each cover 1 KB of consecutive bytes (store q of lane l is copy l & 3
	// of the group's particle q * 8 + l / 4, as in k_expand_sorted); here every lane loads the record it stores itself --
	// the four lanes of a particle ask for the same sector, one L2 request -- so no shuffle transposes the group.
	float4 *dst = pool.vertices + 8ull * (segBase + base);
	uint32_t *vout = pool.vals + segBase + base;
	// the parity read-back (sprReadStreamSlots) turns the arrival index into the slot: bit 31 marks the encoding
	for (uint32_t j = tid; j < n; j += kFinThreads) vout[j] = 0x80000000u | (base + sVal[j]);
	for (uint32_t wbase = warp * 32; wbase < n; wbase += kFinThreads) {
		const uint32_t cnt = min(32u, n - wbase);
		float4 a[4], b[4];
#pragma unroll
		for (uint32_t q = 0; q < 4; q++) {
			const uint32_t p = q * 8 + (lane >> 2);
			a[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); b[q] = a[q];
			if (p < cnt) bk_ld_record_cg(rec + 2ull * sVal[wbase + p], a[q], b[q]);
		}
		float4 *d = dst + 8ull * wbase;
#pragma unroll
		for (uint32_t q = 0; q < 4; q++) {
			const uint32_t p = q * 8 + (lane >> 2);
			if (p < cnt && !(dbg & 4u)) { if (dbg & 16u) bk_st_record_cs(d + 2 * (q * 32 + lane), a[q], b[q]); else bk_st_record_pol(d + 2 * (q * 32 + lane), a[q], b[q], polStream); }
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// k_bucket_splitters: closes the frame, one CTA per segment. An effect the radix path sorted this frame (no table yet,
// or a bucket too large) gets its next table from the sorted slots (PoolPtrs::vals after k_expand_sorted); every table
// then gets its look-up table: cell c covers keys [lo + (c << shift), lo + ((c + 1) << shift)) of the splitter range
// and holds the bucket of its first key.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bucket_splitters(PoolPtrs pool, TablePtrs tab, const uint32_t *segEffects, const BucketSeg *segs,
	const uint32_t *bucketOk, uint32_t *splitters, uint16_t *lut, uint32_t *lutMeta, float camX, float camY, float camZ)
{
	pdl_enter();
	__shared__ uint32_t sSpl[kMaxBins];
	const uint32_t s = blockIdx.x;
	const BucketSeg seg = segs[s];
	if (!seg.nextBins) return;
	const uint32_t g = segEffects[s];
	const unsigned long long N = tab.state[g].aliveCount, nb = seg.nextBins;
	const uint32_t segBase = tab.params[g].slotBase;
	uint32_t *next = splitters + ((size_t)seg.table * 2 + (seg.parity ^ 1u)) * kMaxBins;
	const uint32_t padded = pow2ceil_dev(seg.nextBins);
	const bool fromSlots = !(seg.bins && bucketOk[s]);
	for (uint32_t j = threadIdx.x; j < padded; j += 256) {
		uint32_t v;
		if (fromSlots) {
			v = 0xffffffffu;
			if (j == 0) v = 0;
			else if (j < nb && N) {
				const uint32_t slot = pool.vals[segBase + (uint32_t)(j * N / nb)];
				const float4 p = pool.state[2ull * (segBase + slot)];
				v = depth_key(p.x, p.y, p.z, camX, camY, camZ);
			}
			next[j] = v;
		} else v = next[j];
		sSpl[j] = v;
	}
	__syncthreads();
	const uint32_t lo = sSpl[nb > 1 ? 1 : 0], hi = sSpl[nb - 1];
	const uint32_t range = hi >= lo ? hi - lo : 0u;
	uint32_t shift = 0;
	while ((range >> shift) >= kLutCells) shift++;
	uint16_t *lutOut = lut + ((size_t)seg.table * 2 + (seg.parity ^ 1u)) * (kLutCells + 2);
	for (uint32_t c = threadIdx.x; c <= kLutCells; c += 256) {
		const unsigned long long v64 = (unsigned long long)lo + ((unsigned long long)c << shift);
		const uint32_t v = v64 > 0xffffffffull ? 0xffffffffu : (uint32_t)v64;
		uint32_t pos = 0;
		for (uint32_t st = padded >> 1; st; st >>= 1) if (sSpl[pos + st] <= v) pos += st;
		lutOut[c] = (uint16_t)(c == kLutCells ? nb - 1 : min(pos, (uint32_t)nb - 1u));
	}
	if (threadIdx.x == 0) {
		uint32_t *meta = lutMeta + ((size_t)seg.table * 2 + (seg.parity ^ 1u)) * 4;
		meta[0] = lo; meta[1] = shift;
	}
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static size_t scatter_smem_bytes(uint32_t stride) { return (size_t)stride * (4 + kScWarps * 6); }

void configure_bucket_kernels()
{
	cudaFuncSetAttribute(k_bucket_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem_bytes(kMaxBins));
	cudaFuncSetAttribute(k_bucket_scatter, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
	cudaFuncSetAttribute(k_bucket_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFinSmemBytes);
	cudaFuncSetAttribute(k_bucket_finish, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

uint32_t launch_depth_buckets(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, const BucketArgs &bk,
	const uint32_t *segEffects, uint32_t numBigSegments, const SortTile *tiles, uint32_t numBigTiles, const SortTimingHooks *hooks)
{
	if (!bk.anyBins || !numBigSegments || !numBigTiles) return 0;
	void *tok = nullptr;
#define T_BEGIN(k) do { if (hooks) tok = hooks->begin(hooks->ctx, k); } while (0)
#define T_END(k) do { if (hooks) hooks->end(hooks->ctx, k, tok); } while (0)
	const uint32_t stride = bk.stride;
	static const uint32_t dbg = getenv("SPR_BUCKET_DBG") ? (uint32_t)atoi(getenv("SPR_BUCKET_DBG")) : 0u; // timing experiments only (results are wrong)
	T_BEGIN(5);
	launch_chain(k_bucket_hist, numBigTiles, 256, 0, st, pool, tab, segEffects, tiles, bk.segs, (const uint32_t*)bk.splitters, (const uint16_t*)bk.lut, (const uint32_t*)bk.lutMeta, bk.binId, bk.tileCnt, stride);
	T_END(5);
	T_BEGIN(6);
	launch_chain(k_bucket_scan_tiles, (bk.numChunks * (stride / 32) * 32 + 255) / 256, 256, 0, st, bk.segs, bk.chunks, bk.numChunks,
		(const uint16_t*)bk.tileCnt, bk.tileOff, bk.chunkOff, stride);
	launch_chain(k_bucket_scan_bins, numBigSegments, 1024, 0, st, tab, segEffects, bk.segs, bk.chunkOff, bk.bucketBase, bk.bucketCount, bk.bucketOk, bk.stats, bk.splitters, stride);
	T_END(6);
	T_BEGIN(7);
	launch_chain(k_bucket_scatter, numBigTiles, kScThreads, scatter_smem_bytes(stride), st, pool, tab, segEffects, tiles, bk.segs, (const uint32_t*)bk.bucketOk,
		(const uint16_t*)bk.binId, (const uint32_t*)bk.tileOff, (const uint32_t*)bk.chunkOff, (const uint32_t*)bk.bucketBase, bk.part, stride, dbg);
	T_END(7);
	T_BEGIN(8);
	launch_chain(k_bucket_finish, bk.numJobs, kFinThreads, kFinSmemBytes, st, pool, tab, segEffects, bk.segs, bk.jobs, (const uint32_t*)bk.bucketOk,
		(const uint32_t*)bk.bucketBase, (const uint32_t*)bk.bucketCount, (const float4*)bk.part, bk.splitters, stride, bk.camera[0], bk.camera[1], bk.camera[2], dbg);
	T_END(8);
	return 5;
}

uint32_t launch_bucket_splitters(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, const BucketArgs &bk,
	const uint32_t *segEffects, uint32_t numBigSegments, const SortTimingHooks *hooks)
{
	if (!bk.anyNextBins || !numBigSegments) return 0;
	void *tok = nullptr;
	T_BEGIN(9);
	launch_chain(k_bucket_splitters, numBigSegments, 256, 0, st, pool, tab, segEffects, bk.segs, (const uint32_t*)bk.bucketOk, bk.splitters, bk.lut, bk.lutMeta,
		bk.camera[0], bk.camera[1], bk.camera[2]);
	T_END(9);
#undef T_BEGIN
#undef T_END
	return 1;
}

} // namespace spr
