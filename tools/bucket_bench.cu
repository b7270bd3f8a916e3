// Micro-benchmark + self-check of the bucketed depth sort (tools/bucket_sort.cuh) on a C3-shaped input:
// 10M records in a box around an emitter, depth keys ~enc(|cam - p|^2).
//   k_scatter : model of the integration pass with the bucket scatter (read 32 B, write 32 B in place, key,
//               atomic slot reservation in the bucket, 32 B record + key + slot to the bucket)
//   k_prep    : per-bucket offsets (exclusive scan), overflow check
//   k_finish  : bucket_sort_expand per bucket
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false tools/bucket_bench.cu -o tools/bucket_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <random>

#include "bucket_sort.cuh"

using namespace spr;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__host__ __device__ inline uint32_t enc_float_bits(uint32_t u) { return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }

__device__ inline void ld_rec(const float4 *p, float4 &a, float4 &b)
{
	asm("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ inline void st_rec(float4 *p, const float4 &a, const float4 &b)
{
	asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
		:: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

struct Map { uint32_t lo, scale, numBins; };

// MODE 0 integrate only | 1 + atomics (positions to a coalesced array) | 2 + record scatter at precomputed positions (no atomics)
//      3 atomics + record scatter | 4 = 3 + (key, slot) as one 8-byte store | 5 = 3 + key and slot as two 4-byte stores | 6 = 1 with RED (no return value)
template <int MODE>
__global__ void __launch_bounds__(256, 3) k_scatter(float4 *state, uint32_t n, float cx, float cy, float cz, Map map,
	uint32_t *cursor, uint32_t cursorStride, float4 *bRecs, uint32_t *bKeys, uint32_t *bSlots, uint2 *bPairs, uint32_t *slotKeys, uint32_t *posArr)
{
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t granule = blockIdx.x * 8 + warp;
	float4 a[4], b[4];
	uint32_t slot0 = granule * 128 + lane;
	float4 *p = state + 2ull * slot0;
#pragma unroll
	for (int r = 0; r < 4; r++) if (slot0 + r * 32 < n) ld_rec(p + r * 64, a[r], b[r]);
	uint32_t key[4], pos[4];
#pragma unroll
	for (int r = 0; r < 4; r++) {
		pos[r] = 0xffffffffu; key[r] = 0;
		if (slot0 + r * 32 >= n) continue;
		a[r].x += b[r].x * 0.0f; a[r].y += b[r].y * 0.0f; a[r].z += b[r].z * 0.0f;
		st_rec(p + r * 64, a[r], b[r]);
		float dx = cx - a[r].x, dy = cy - a[r].y, dz = cz - a[r].z;
		key[r] = ~enc_float_bits(__float_as_uint(dx*dx + dy*dy + dz*dz));
		slotKeys[slot0 + r * 32] = key[r];
		if (MODE == 1 || MODE >= 3) {
			uint32_t bin = bucket_of(key[r], map.lo, map.scale, map.numBins);
			if (MODE == 6) { atomicAdd(&cursor[bin * cursorStride], 1u); continue; }
			uint32_t i = atomicAdd(&cursor[bin * cursorStride], 1u);
			if (i < kBucketCap) pos[r] = bin * kBucketCap + i;
		}
		if (MODE == 2) pos[r] = posArr[slot0 + r * 32];
	}
#pragma unroll
	for (int r = 0; r < 4; r++) {
		if (pos[r] == 0xffffffffu) continue;
		if (MODE == 1) posArr[slot0 + r * 32] = pos[r];
		if (MODE >= 2 && MODE <= 5) st_rec(bRecs + 2ull * pos[r], a[r], b[r]);
		if (MODE == 4) bPairs[pos[r]] = make_uint2(key[r], slot0 + r * 32);
		if (MODE == 5) { bKeys[pos[r]] = key[r]; bSlots[pos[r]] = slot0 + r * 32; }
	}
}

__global__ void __launch_bounds__(256) k_prep(const uint32_t *cursor, uint32_t cursorStride, uint32_t numBins, uint32_t *offsets, uint32_t *flags)
{
	__shared__ uint32_t sWarp[8];
	__shared__ uint32_t sRun;
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) sRun = 0;
	__syncthreads();
	bool over = false;
	for (uint32_t base = 0; base < numBins; base += 256 * 8) {
		uint32_t v[8], sum = 0;
#pragma unroll
		for (int i = 0; i < 8; i++) { uint32_t b = base + tid * 8 + i; v[i] = b < numBins ? cursor[b * cursorStride] : 0; over |= v[i] > kBucketCap; sum += v[i]; }
		uint32_t incl = sum;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += o; }
		if (lane == 31) sWarp[warp] = incl;
		__syncthreads();
		uint32_t run = sRun;
		for (uint32_t w = 0; w < warp; w++) run += sWarp[w];
		run += incl - sum;
#pragma unroll
		for (int i = 0; i < 8; i++) { uint32_t b = base + tid * 8 + i; if (b < numBins) offsets[b] = run; run += v[i]; }
		__syncthreads();
		if (tid == 255) sRun = run;
		__syncthreads();
	}
	if (over) flags[0] = 1;
	if (tid == 0) flags[1] = sRun;
}

__global__ void __launch_bounds__(kBucketWarpsPerCta * 32, 4) k_finish(uint32_t *cursor, uint32_t cursorStride, uint32_t numBins, const uint32_t *offsets, const float4 *bRecs, const uint2 *bPairs,
	float4 *verts, uint32_t *vals)
{
	__shared__ BucketSmem smem[kBucketWarpsPerCta];
	const uint32_t warp = threadIdx.x >> 5;
	uint32_t bin = blockIdx.x * kBucketWarpsPerCta + warp;
	if (bin >= numBins) return;
	uint32_t n = cursor[bin * cursorStride];
	if (n == 0) return;
	__syncwarp();
	if ((threadIdx.x & 31) == 0) cursor[bin * cursorStride] = 0;
	uint32_t off = offsets[bin];
	uint64_t rb = (uint64_t)bin * kBucketCap;
	bucket_sort_expand(smem[warp], bPairs + rb, bRecs + 2 * rb, n, verts + 8ull * off, vals + off);
}

// gather-only / identity model of the finishing kernel (what its memory phase alone costs)
__global__ void __launch_bounds__(256) k_finish_copy(const uint32_t *cursor, uint32_t cursorStride, const uint32_t *offsets, const float4 *bRecs, float4 *verts)
{
	uint32_t bin = blockIdx.x, n = cursor[bin * cursorStride], off = offsets[bin];
	const float4 *src = bRecs + 2ull * bin * kBucketCap;
	for (uint32_t i = threadIdx.x; i < n; i += 64) {
		float4 a, b;
		bk_ld_record(src + 2ull * i, a, b);
		float4 *d = verts + 8ull * (off + i);
		bk_st_record(d, a, b); bk_st_record(d + 2, a, b); bk_st_record(d + 4, a, b); bk_st_record(d + 6, a, b);
	}
}

template <int MODE>
static float time_scatter(float4 *state, uint32_t n, Map map, uint32_t *cursor, uint32_t stride, size_t cursorBytes, float4 *bRecs, uint32_t *bKeys, uint32_t *bSlots,
	uint2 *bPairs, uint32_t *slotKeys, uint32_t *posArr)
{
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	uint32_t grid = (n + 1023) / 1024;
	float total = 0;
	const int reps = 6;
	for (int it = 0; it < reps + 2; it++) {
		CK(cudaMemsetAsync(cursor, 0, cursorBytes));
		cudaEventRecord(e0);
		k_scatter<MODE><<<grid, 256>>>(state, n, 16, 20, -30, map, cursor, stride, bRecs, bKeys, bSlots, bPairs, slotKeys, posArr);
		cudaEventRecord(e1);
		CK(cudaDeviceSynchronize());
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		if (it >= 2) total += ms / reps;
	}
	return total;
}

int main(int argc, char **argv)
{
	const uint32_t n = argc > 1 ? (uint32_t)atoi(argv[1]) : 10000000u;
	const uint32_t dupEvery = argc > 2 ? (uint32_t)atoi(argv[2]) : 0u; // > 0: every dupEvery-th record duplicates its predecessor's position (ties)
	std::vector<float> h((size_t)n * 8);
	std::mt19937 rng(7);
	std::uniform_real_distribution<float> U(0.0f, 1.0f);
	for (uint32_t i = 0; i < n; i++) {
		float *r = &h[(size_t)i * 8];
		r[0] = 1.0f + (U(rng) - 0.5f) * 2.0f; r[1] = 2.0f + (U(rng) - 0.5f) * 1.0f; r[2] = 3.0f + (U(rng) - 0.5f) * 2.0f; r[3] = 0.5f;
		r[4] = U(rng); r[5] = U(rng); r[6] = U(rng); r[7] = (float)i;
		if (dupEvery && i && i % dupEvery == 0) { r[0] = r[-8]; r[1] = r[-7]; r[2] = r[-6]; }
	}
	const float cx = 16, cy = 20, cz = -30;
	uint32_t kmin = 0xffffffffu, kmax = 0;
	for (uint32_t i = 0; i < n; i++) {
		float dx = cx - h[(size_t)i*8], dy = cy - h[(size_t)i*8+1], dz = cz - h[(size_t)i*8+2];
		float d = dx*dx; d += dy*dy; d += dz*dz;
		uint32_t u; memcpy(&u, &d, 4);
		uint32_t k = ~enc_float_bits(u);
		kmin = std::min(kmin, k); kmax = std::max(kmax, k);
	}
	float4 *state, *bRecs, *verts; uint32_t *cursor, *offsets, *bKeys, *bSlots, *vals, *slotKeys, *flags, *posArr; uint2 *bPairs;
	for (uint32_t perBin : { 160u, 224u }) {
		uint32_t numBins = std::max(1u, n / perBin);
		uint64_t width = (uint64_t)kmax - kmin + 1, margin = width / 16 + 1;
		uint64_t lo = kmin > margin ? kmin - margin : 0, hi = std::min<uint64_t>(0xffffffffull, (uint64_t)kmax + margin);
		uint64_t W = hi - lo + 1;
		uint64_t scale = numBins > 1 ? ((uint64_t)numBins << 32) / W : 0;
		if (scale > 0xffffffffull) scale = 0xffffffffull;
		Map map = { (uint32_t)lo, (uint32_t)scale, numBins };
		const size_t cursorBytes = (size_t)numBins * 32 * 4;
		CK(cudaMalloc(&state, (size_t)n * 32)); CK(cudaMalloc(&bRecs, (size_t)numBins * kBucketCap * 32)); CK(cudaMalloc(&verts, (size_t)n * 128));
		CK(cudaMalloc(&cursor, cursorBytes)); CK(cudaMalloc(&offsets, numBins * 4)); CK(cudaMalloc(&bKeys, (size_t)numBins * kBucketCap * 4));
		CK(cudaMalloc(&bSlots, (size_t)numBins * kBucketCap * 4)); CK(cudaMalloc(&bPairs, (size_t)numBins * kBucketCap * 8));
		CK(cudaMalloc(&vals, (size_t)n * 4)); CK(cudaMalloc(&slotKeys, (size_t)n * 4)); CK(cudaMalloc(&flags, 8)); CK(cudaMalloc(&posArr, (size_t)n * 4));
		CK(cudaMemcpy(state, h.data(), (size_t)n * 32, cudaMemcpyHostToDevice));
		CK(cudaMemset(flags, 0, 8));
		CK(cudaFuncSetAttribute(k_finish, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
		int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_finish, kBucketWarpsPerCta * 32, 0));
		const uint32_t STRIDE = kCursorStride;
		printf("n=%u bins=%u (avg %u/bucket), finish occupancy %d CTAs/SM, smem %zu\n", n, numBins, n / numBins, occ, sizeof(BucketSmem));
#define TS(MODE, stride) time_scatter<MODE>(state, n, map, cursor, stride, cursorBytes, bRecs, bKeys, bSlots, bPairs, slotKeys, posArr)
		printf("  scatter kernel: integrate only %.3f | +atomics %.3f (stride 32: %.3f, stride 8: %.3f) | +record scatter, no atomics %.3f | atomics(s32)+records %.3f | +pairs(8B) s8 %.3f, s32 %.3f | +key,slot(4B+4B) s32 %.3f ms\n",
			TS(0, 1), TS(1, 1), TS(1, 32), TS(1, 8), TS(2, 1), TS(3, 32), TS(4, 8), TS(4, 32), TS(5, 32));
		// the real sequence: scatter (pairs) -> prep -> finish
		cudaEvent_t e[4]; for (auto &x : e) cudaEventCreate(&x);
		float t[3] = { 0, 0, 0 }, tcopy = 0;
		const int reps = 10;
		uint32_t grid = (n + 1023) / 1024;
		CK(cudaMemset(cursor, 0, cursorBytes));
		for (int it = 0; it < reps + 2; it++) {
			cudaEventRecord(e[0]);
			k_scatter<4><<<grid, 256>>>(state, n, cx, cy, cz, map, cursor, STRIDE, bRecs, bKeys, bSlots, bPairs, slotKeys, posArr);
			cudaEventRecord(e[1]);
			k_prep<<<1, 256>>>(cursor, STRIDE, numBins, offsets, flags);
			cudaEventRecord(e[2]);
			if (it == reps + 1) {
				cudaEvent_t c0, c1; cudaEventCreate(&c0); cudaEventCreate(&c1);
				for (int k = 0; k < 3; k++) { if (k == 1) cudaEventRecord(c0); k_finish_copy<<<numBins, 64>>>(cursor, STRIDE, offsets, bRecs, verts); }
				cudaEventRecord(c1); CK(cudaDeviceSynchronize());
				cudaEventElapsedTime(&tcopy, c0, c1); tcopy /= 2;
				cudaEventRecord(e[2]);
			}
			k_finish<<<(numBins + kBucketWarpsPerCta - 1) / kBucketWarpsPerCta, kBucketWarpsPerCta * 32>>>(cursor, STRIDE, numBins, offsets, bRecs, bPairs, verts, vals);
			cudaEventRecord(e[3]);
			CK(cudaDeviceSynchronize());
			if (it >= 2) for (int k = 0; k < 3; k++) { float ms; cudaEventElapsedTime(&ms, e[k], e[k + 1]); t[k] += ms / reps; }
		}
		uint32_t hf[2]; CK(cudaMemcpy(hf, flags, 8, cudaMemcpyDeviceToHost));
		printf("  sequence: scatter %.3f ms, prep %.3f ms, finish %.3f ms (its memory phase alone, in arrival order: %.3f ms) | total %.3f ms | overflow=%u total=%u\n",
			t[0], t[1], t[2], tcopy, t[0] + t[1] + t[2], hf[0], hf[1]);
		// ---- check against a stable sort of the device keys
		std::vector<uint32_t> dk(n), dv(n);
		CK(cudaMemcpy(dk.data(), slotKeys, (size_t)n * 4, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(dv.data(), vals, (size_t)n * 4, cudaMemcpyDeviceToHost));
		std::vector<uint32_t> order(n);
		for (uint32_t i = 0; i < n; i++) order[i] = i;
		std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return dk[x] < dk[y]; });
		size_t bad = 0;
		for (uint32_t i = 0; i < n; i++) if (order[i] != dv[i]) { if (bad < 5) printf("  mismatch at %u: want slot %u got %u\n", i, order[i], dv[i]); bad++; }
		std::vector<float> hv((size_t)4096 * 32);
		CK(cudaMemcpy(hv.data(), verts, hv.size() * 4, cudaMemcpyDeviceToHost));
		size_t badv = 0;
		for (uint32_t i = 0; i < 4096 && i < n; i++) for (int c = 0; c < 4; c++) if (memcmp(&hv[(size_t)i * 32 + c * 8], &h[(size_t)order[i] * 8], 32) != 0) badv++;
		printf("  order mismatches: %zu, vertex mismatches (first 4096): %zu\n", bad, badv);
		cudaFree(state); cudaFree(bRecs); cudaFree(verts); cudaFree(cursor); cudaFree(offsets); cudaFree(bKeys); cudaFree(bSlots); cudaFree(bPairs);
		cudaFree(vals); cudaFree(slotKeys); cudaFree(flags); cudaFree(posArr);
	}
	return 0;
}
