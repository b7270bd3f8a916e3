// Micro-benchmark: the sorted expansion (random 32-byte gather + 4x store, 128 contiguous bytes per particle) with the
// stores issued by the TMA unit instead of the LSU.
//   variant 0  product pattern: registers -> shuffle transpose -> 4 x st.v8 per lane (k_expand_sorted)
//   variant 1  records -> compact shared memory (32 B each) -> 4 x cp.async.bulk.tensor.3d per 128 records: the vertex
//              stream seen as [particle][copy 0..3][8 floats], box [128][1][8], one store per copy
//   variant 2  records -> shared memory REPLICATED 4x (128 B each) -> one 1-D cp.async.bulk of 16 KB per 128 records
//   variant 3  as 1, but the gather itself is cp.async (global -> shared, no registers), then the tensor stores
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_store_bench tools/tma_store_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ inline void ld_rec(const float4 *p, float4 &a, float4 &b)
{
	asm volatile("ld.global.L2::64B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p));
}
__device__ inline void st_rec(float4 *p, const float4 &a, const float4 &b)
{
	asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(a.x),"f"(a.y),"f"(a.z),"f"(a.w),"f"(b.x),"f"(b.y),"f"(b.z),"f"(b.w) : "memory");
}
__device__ inline uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline void tma_store_3d(const CUtensorMap *map, const void *smem, int c0, int c1, int c2)
{
	asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
		:: "l"((uint64_t)map), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ inline void bulk_store_1d(void *gmem, const void *smem, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gmem), "r"(smem_u32(smem)), "r"(bytes) : "memory");
}
__device__ inline void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ inline void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ inline void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// variant 0: the product kernel's pattern
__global__ void __launch_bounds__(256) k_lsu(const float4 *state, const uint32_t *idx, float4 *out, uint32_t n)
{
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t base = blockIdx.x * 1024;
	uint32_t slot[4]; float4 a[4], b[4];
#pragma unroll
	for (int u = 0; u < 4; u++) { uint32_t i = base + u * 256 + threadIdx.x; slot[u] = i < n ? idx[i] : 0xffffffffu; }
#pragma unroll
	for (int u = 0; u < 4; u++) if (slot[u] != 0xffffffffu) ld_rec(state + 2ull * slot[u], a[u], b[u]);
#pragma unroll
	for (int u = 0; u < 4; u++) {
		const uint32_t wbase = base + u * 256 + warp * 32;
		if (wbase >= n) continue;
		const uint32_t cnt = min(32u, n - wbase);
		float4 *d = out + 8ull * wbase;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const uint32_t p = j * 8 + (lane >> 2);
			float4 x, y;
			x.x = __shfl_sync(0xffffffffu, a[u].x, p); x.y = __shfl_sync(0xffffffffu, a[u].y, p); x.z = __shfl_sync(0xffffffffu, a[u].z, p); x.w = __shfl_sync(0xffffffffu, a[u].w, p);
			y.x = __shfl_sync(0xffffffffu, b[u].x, p); y.y = __shfl_sync(0xffffffffu, b[u].y, p); y.z = __shfl_sync(0xffffffffu, b[u].z, p); y.w = __shfl_sync(0xffffffffu, b[u].w, p);
			if (p < cnt) st_rec(d + 2 * (j * 32 + lane), x, y);
		}
	}
}

// variant 1 / 3: WARPS warps per CTA, each owns 128 consecutive ranks; compact staging (4 KB per warp)
template <int WARPS, bool ASYNC_GATHER>
__global__ void __launch_bounds__(WARPS * 32) k_tma_tensor(const float4 *state, const uint32_t *idx, float4 *out, uint32_t n, const __grid_constant__ CUtensorMap map)
{
	__shared__ __align__(128) float4 stage[WARPS][256];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t base = (blockIdx.x * WARPS + warp) * 128;
	if (base >= n) return;
	uint32_t slot[4]; float4 a[4], b[4];
#pragma unroll
	for (int u = 0; u < 4; u++) { uint32_t i = base + u * 32 + lane; slot[u] = i < n ? idx[i] : 0xffffffffu; }
	if (ASYNC_GATHER) {
#pragma unroll
		for (int u = 0; u < 4; u++) if (slot[u] != 0xffffffffu) {
			const uint32_t dst = smem_u32(&stage[warp][2 * (u * 32 + lane)]);
			const float4 *src = state + 2ull * slot[u];
			asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
			asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" :: "r"(dst + 16), "l"(src + 1) : "memory");
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
		asm volatile("cp.async.wait_group 0;" ::: "memory");
	} else {
#pragma unroll
		for (int u = 0; u < 4; u++) if (slot[u] != 0xffffffffu) ld_rec(state + 2ull * slot[u], a[u], b[u]);
#pragma unroll
		for (int u = 0; u < 4; u++) { stage[warp][2 * (u * 32 + lane)] = a[u]; stage[warp][2 * (u * 32 + lane) + 1] = b[u]; }
	}
	if (base + 128 <= n) {
		fence_async_smem();
		__syncwarp();
		if (lane == 0) {
#pragma unroll
			for (int c = 0; c < 4; c++) tma_store_3d(&map, &stage[warp][0], 0, c, (int)base);
			bulk_commit();
			bulk_wait_read();
		}
	} else { // ragged tail: plain stores
		__syncwarp();
		const uint32_t cnt = n - base;
		for (uint32_t q = lane; q < cnt * 8; q += 32) out[8ull * base + q] = stage[warp][((q >> 3) << 1) | (q & 1u)];
	}
}

// variant 2: replicated staging (16 KB per warp), one contiguous bulk store per warp
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_bulk_1d(const float4 *state, const uint32_t *idx, float4 *out, uint32_t n)
{
	extern __shared__ __align__(128) float4 stageDyn[];
	float4 *stage = stageDyn + (threadIdx.x >> 5) * 1024;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t base = (blockIdx.x * WARPS + warp) * 128;
	if (base >= n) return;
	uint32_t slot[4]; float4 a[4], b[4];
#pragma unroll
	for (int u = 0; u < 4; u++) { uint32_t i = base + u * 32 + lane; slot[u] = i < n ? idx[i] : 0xffffffffu; }
#pragma unroll
	for (int u = 0; u < 4; u++) if (slot[u] != 0xffffffffu) ld_rec(state + 2ull * slot[u], a[u], b[u]);
#pragma unroll
	for (int u = 0; u < 4; u++) {
		float4 *d = stage + 8 * (u * 32 + lane);
#pragma unroll
		for (int c = 0; c < 4; c++) { d[2 * c] = a[u]; d[2 * c + 1] = b[u]; }
	}
	const uint32_t cnt = min(128u, n - base);
	fence_async_smem();
	__syncwarp();
	if (lane == 0) {
		bulk_store_1d(out + 8ull * base, stage, cnt * 128);
		bulk_commit();
		bulk_wait_read();
	}
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
	CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(float4 *out, uint64_t n, CUtensorMapL2promotion promo)
{
	void *fn = nullptr; cudaDriverEntryPointQueryResult qr;
	CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
	CUtensorMap m;
	cuuint64_t dims[3] = { 8, 4, n };
	cuuint64_t strides[2] = { 32, 128 };
	cuuint32_t box[3] = { 8, 1, 128 };
	cuuint32_t es[3] = { 1, 1, 1 };
	CUresult r = ((EncodeFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
	return m;
}

template <typename F>
static float timeit(F f)
{
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int i = 0; i < 2; i++) f();
	cudaEventRecord(e0);
	for (int i = 0; i < 10; i++) f();
	cudaEventRecord(e1); CK(cudaDeviceSynchronize());
	CK(cudaGetLastError());
	float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 10;
}

static size_t count_bad(const float4 *out, uint32_t n, const std::vector<uint32_t> &h)
{
	std::vector<float> o((size_t)n * 32);
	CK(cudaMemcpy(o.data(), out, (size_t)n * 128, cudaMemcpyDeviceToHost));
	size_t bad = 0;
	for (uint32_t i = 0; i < n; i++) for (int c = 0; c < 4; c++) if (o[(size_t)i * 32 + c * 8] != (float)h[i] || o[(size_t)i * 32 + c * 8 + 7] != (float)(h[i] & 1023)) bad++;
	return bad;
}

int main()
{
	CK(cudaFuncSetAttribute(k_bulk_1d<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384));
	CK(cudaFuncSetAttribute(k_bulk_1d<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384));
	for (uint32_t n : { 10000000u, 1048640u }) {
		std::vector<uint32_t> h(n);
		for (uint32_t i = 0; i < n; i++) h[i] = i;
		std::mt19937 rng(1); std::shuffle(h.begin(), h.end(), rng);
		float4 *state, *out; uint32_t *idx;
		CK(cudaMalloc(&state, (size_t)n * 32)); CK(cudaMalloc(&out, (size_t)n * 128)); CK(cudaMalloc(&idx, (size_t)n * 4));
		{
			std::vector<float> s((size_t)n * 8, 0.0f);
			for (uint32_t i = 0; i < n; i++) { s[(size_t)i * 8] = (float)i; s[(size_t)i * 8 + 7] = (float)(i & 1023); }
			CK(cudaMemcpy(state, s.data(), (size_t)n * 32, cudaMemcpyHostToDevice));
		}
		CUtensorMap map = make_map(out, n, CU_TENSOR_MAP_L2_PROMOTION_NONE);
		CUtensorMap map256 = make_map(out, n, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
		for (int order = 0; order < 2; order++) {
			if (order == 1) for (uint32_t i = 0; i < n; i++) h[i] = i;
			CK(cudaMemcpy(idx, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
			printf("n = %u (%s order)\n", n, order ? "identity" : "random");
			const uint32_t m = n < 2000000u ? n - 7 : n; // ragged tail on purpose for the small case
			auto v0 = [&] { k_lsu<<<(m + 1023) / 1024, 256>>>(state, idx, out, m); };
			auto v1a = [&] { k_tma_tensor<4, false><<<(m + 511) / 512, 128>>>(state, idx, out, m, map); };
			auto v1b = [&] { k_tma_tensor<8, false><<<(m + 1023) / 1024, 256>>>(state, idx, out, m, map); };
			auto v1c = [&] { k_tma_tensor<8, false><<<(m + 1023) / 1024, 256>>>(state, idx, out, m, map256); };
			auto v2a = [&] { k_bulk_1d<4><<<(m + 511) / 512, 128, 4 * 16384>>>(state, idx, out, m); };
			auto v2b = [&] { k_bulk_1d<8><<<(m + 1023) / 1024, 256, 8 * 16384>>>(state, idx, out, m); };
			auto v3a = [&] { k_tma_tensor<4, true><<<(m + 511) / 512, 128>>>(state, idx, out, m, map); };
			auto v3b = [&] { k_tma_tensor<8, true><<<(m + 1023) / 1024, 256>>>(state, idx, out, m, map); };
			if (n < 2000000u && order == 0) {
				int k = 0;
				for (auto *f : { (void*)0 }) (void)f;
				auto chk = [&](const char *name, auto fn) { CK(cudaMemset(out, 0xff, (size_t)n * 128)); fn(); CK(cudaDeviceSynchronize()); printf("  check %-28s: %zu bad copies\n", name, count_bad(out, m, h)); k++; };
				chk("lsu", v0); chk("tma tensor 4w", v1a); chk("tma tensor 8w", v1b); chk("bulk 1d 4w", v2a); chk("bulk 1d 8w", v2b); chk("cp.async + tma 4w", v3a); chk("cp.async + tma 8w", v3b);
			}
			printf("  LSU shuffle transpose (product)           : %.4f ms\n", timeit(v0));
			printf("  TMA tensor store, compact smem, 4 warps   : %.4f ms\n", timeit(v1a));
			printf("  TMA tensor store, compact smem, 8 warps   : %.4f ms\n", timeit(v1b));
			printf("  same, L2 promotion 256B                   : %.4f ms\n", timeit(v1c));
			printf("  1-D bulk store, replicated smem, 4 warps  : %.4f ms\n", timeit(v2a));
			printf("  1-D bulk store, replicated smem, 8 warps  : %.4f ms\n", timeit(v2b));
			printf("  cp.async gather + TMA tensor store, 4 w   : %.4f ms\n", timeit(v3a));
			printf("  cp.async gather + TMA tensor store, 8 w   : %.4f ms\n", timeit(v3b));
		}
		cudaFree(state); cudaFree(out); cudaFree(idx);
	}
	return 0;
}
