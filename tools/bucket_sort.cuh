// bucket_sort.cuh -- EXPERIMENT (not part of the product): the shared-memory half of a bucketed depth sort,
// measured by tools/bucket_bench.cu and rejected (profiles/r01_bucket_microbench.txt).
//
// The integration pass drops every live record into one of the effect's depth buckets (a monotone
// map of the depth key, so bucket b holds only keys <= those of bucket b+1). One CTA then finishes
// a bucket -- one WARP per bucket, so that no phase of it ever waits on a block barrier --: it sorts the bucket's (key, slot) pairs in shared memory (counting sort over the narrow
// key range of the bucket; the records arrive in atomic order, so the order is rebuilt from the
// pairs alone: depth descending, ties by ascending slot, the BillboardOrder idiom of
// src/client/BillboardSystem.cpp:41-51), and writes the
// records four times each in sorted order (ParticleSystem.cpp:607-665). The records of a bucket are
// contiguous, so every DRAM line the gather touches is used completely -- unlike the random
// 32-byte gather of the sort-indices-then-gather path, which pays a 128-byte line per record.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace spr {

static const uint32_t kBucketCap = 512;         // records a bucket can hold
static const uint32_t kBucketItems = kBucketCap / 32; // 16 keys per lane at most
static const uint32_t kBucketHistBins = 256;    // counting-sort bins of one bucket
static const uint32_t kBucketWarpsPerCta = 8;
static const uint32_t kCursorStride = 32;       // one bucket counter per 128-byte line (atomics to one line serialise in L2)

// Shared memory of ONE warp (a warp finishes a bucket on its own: no block barriers anywhere).
struct BucketSmem {
	uint32_t hist[kBucketHistBins];         // counts -> bin starts -> (after the scatter) bin ends
	uint32_t slot[kBucketCap];              // slot of the i-th arrival
	uint16_t idx[2][kBucketCap];            // arrival index of the record at each sorted position
};

__device__ __forceinline__ void bk_ld_record(const float4 *p, float4 &a, float4 &b)
{
	asm("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ void bk_st_record(float4 *p, const float4 &a, const float4 &b)
{
	asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
		:: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// One WARP sorts bucket [0, n) (n <= kBucketCap) by (key, slot) and expands it: outVerts gets 4 x 32 B per
// record, outVals the slot of each. gPairs[i] = (key, slot) of the i-th arrival, gRecs its record.
//
// The keys of one bucket span a narrow range (the bucket map is linear in the key), so the sort is a
// counting sort on (key - kmin) >> sh with sh = 0 in the common case: one histogram of shared-memory
// atomics, one scan, one scatter, all from registers. A float depth has few distinct values over a
// narrow range, so several particles per key value are the norm in a big effect: elements that share
// a bin (equal keys, or close keys when sh > 0) are ordered exactly by counting inside the bin.
__device__ inline void bucket_sort_expand(BucketSmem &s, const uint2 *gPairs, const float4 *gRecs,
	uint32_t n, float4 *outVerts, uint32_t *outVals)
{
	const uint32_t lane = threadIdx.x & 31;
	uint32_t key[kBucketItems];
	uint32_t mn = 0xffffffffu, mx = 0;
#pragma unroll
	for (uint32_t u = 0; u < kBucketItems; u++) {
		uint32_t i = u * 32 + lane;
		key[u] = 0;
		if (i < n) { uint2 p = gPairs[i]; key[u] = p.x; s.slot[i] = p.y; mn = min(mn, p.x); mx = max(mx, p.x); }
	}
#pragma unroll
	for (uint32_t j = 0; j < kBucketHistBins / 32; j++) s.hist[j * 32 + lane] = 0;
	const uint32_t kmin = __reduce_min_sync(0xffffffffu, mn);
	const uint32_t range = __reduce_max_sync(0xffffffffu, mx) - kmin;
	const uint32_t sh = range < kBucketHistBins ? 0u : (32u - (uint32_t)__clz(range)) - 8u; // (range >> sh) < 256
	__syncwarp();
#pragma unroll
	for (uint32_t u = 0; u < kBucketItems; u++) if (u * 32 + lane < n) atomicAdd(&s.hist[(key[u] - kmin) >> sh], 1u);
	__syncwarp();
	bool multi = false;
	{
		// exclusive scan of the 256 bins: 8 consecutive bins per lane
		uint32_t c[8], sum = 0;
#pragma unroll
		for (int j = 0; j < 8; j++) { c[j] = s.hist[lane * 8 + j]; sum += c[j]; multi |= c[j] > 1; }
		uint32_t incl = sum;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += o; }
		uint32_t run = incl - sum;
		__syncwarp();
#pragma unroll
		for (int j = 0; j < 8; j++) { s.hist[lane * 8 + j] = run; run += c[j]; }
	}
	multi = __any_sync(0xffffffffu, multi);
	__syncwarp();
#pragma unroll
	for (uint32_t u = 0; u < kBucketItems; u++) {
		uint32_t i = u * 32 + lane;
		if (i < n) s.idx[0][atomicAdd(&s.hist[(key[u] - kmin) >> sh], 1u)] = (uint16_t)i;
	}
	__syncwarp();
	const uint16_t *I = s.idx[0];
	if (multi) {
		// bins with several elements: exact order by (key, slot), rank by counting (hist[v] is now the end of bin v)
		uint16_t *J = s.idx[1];
#pragma unroll
		for (uint32_t u = 0; u < kBucketItems; u++) {
			uint32_t i = u * 32 + lane;
			if (i >= n) continue;
			uint32_t v = (key[u] - kmin) >> sh;
			uint32_t b = s.hist[v], a = v ? s.hist[v - 1] : 0u;
			uint32_t cnt = 0;
			if (b - a > 1) {
				const uint32_t mySlot = s.slot[i];
				if (sh == 0) {
					for (uint32_t j = a; j < b; j++) cnt += s.slot[I[j]] < mySlot ? 1u : 0u;
				} else {
					for (uint32_t j = a; j < b; j++) {
						uint32_t o = I[j], ok = gPairs[o].x;
						cnt += (ok < key[u] || (ok == key[u] && s.slot[o] < mySlot)) ? 1u : 0u;
					}
				}
			}
			J[a + cnt] = (uint16_t)i;
		}
		__syncwarp();
		I = J;
	}
	// ---- expansion in sorted order: one lane per particle, four gathers in flight per lane
	for (uint32_t base = 0; base < n; base += 128) {
		uint32_t id[4];
		float4 a[4], b[4];
#pragma unroll
		for (int u = 0; u < 4; u++) {
			uint32_t i = base + u * 32 + lane;
			id[u] = i < n ? (uint32_t)I[i] : 0xffffffffu;
		}
#pragma unroll
		for (int u = 0; u < 4; u++) if (id[u] != 0xffffffffu) bk_ld_record(gRecs + 2ull * id[u], a[u], b[u]);
#pragma unroll
		for (int u = 0; u < 4; u++) {
			if (id[u] == 0xffffffffu) continue;
			uint32_t i = base + u * 32 + lane;
			float4 *d = outVerts + 8ull * i;
			bk_st_record(d, a[u], b[u]);
			bk_st_record(d + 2, a[u], b[u]);
			bk_st_record(d + 4, a[u], b[u]);
			bk_st_record(d + 6, a[u], b[u]);
			outVals[i] = s.slot[id[u]];
		}
	}
}

// Monotone map of a depth key to a bucket: floor((key - lo) * scale / 2^32), clamped at both ends.
__device__ __forceinline__ uint32_t bucket_of(uint32_t key, uint32_t lo, uint32_t scale, uint32_t numBins)
{
	uint32_t rel = key > lo ? key - lo : 0u;
	uint32_t b = __umulhi(rel, scale);
	return b < numBins ? b : numBins - 1;
}

} // namespace spr
