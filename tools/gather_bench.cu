// Micro-benchmark: random 32-byte record gather (the sorted expansion's access pattern) under
// different L2 fetch-granularity limits and load flavours. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE>
__device__ inline void ld32(const float4 *p, float4 &a, float4 &b)
{
	if (MODE == 0) asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p));
	if (MODE == 1) asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p));
	if (MODE == 2) asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p));
	if (MODE == 3) { a = __ldg(p); b = __ldg(p + 1); }
	if (MODE == 4) { asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w) : "l"(p));
	                 asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p + 1)); }
	if (MODE == 5) asm volatile("ld.global.lu.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p));
	// L2 prefetch-size hints (the smallest is 64 B) and eviction priorities
	if (MODE == 6) asm volatile("ld.global.L2::64B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p));
	if (MODE == 7) { asm volatile("ld.global.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w) : "l"(p));
	                 asm volatile("ld.global.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p + 1)); }
	if (MODE == 8) asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p));
	if (MODE == 9) { asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w) : "l"(p));
	                 asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p + 1)); }
}

template <int MODE, bool WRITE4>
__global__ void __launch_bounds__(256) k_gather(const float4 *state, const uint32_t *idx, float4 *out, uint32_t n)
{
	uint32_t i = blockIdx.x * 1024 + threadIdx.x;
#pragma unroll
	for (int u = 0; u < 4; u++, i += 256) {
		if (i >= n) return;
		float4 a, b;
		ld32<MODE>(state + 2ull * idx[i], a, b);
		float4 *d = out + (WRITE4 ? 8ull : 2ull) * i;
		asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(d), "f"(a.x),"f"(a.y),"f"(a.z),"f"(a.w),"f"(b.x),"f"(b.y),"f"(b.z),"f"(b.w) : "memory");
		if (WRITE4) {
			asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(d + 2), "f"(a.x),"f"(a.y),"f"(a.z),"f"(a.w),"f"(b.x),"f"(b.y),"f"(b.z),"f"(b.w) : "memory");
			asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(d + 4), "f"(a.x),"f"(a.y),"f"(a.z),"f"(a.w),"f"(b.x),"f"(b.y),"f"(b.z),"f"(b.w) : "memory");
			asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(d + 6), "f"(a.x),"f"(a.y),"f"(a.z),"f"(a.w),"f"(b.x),"f"(b.y),"f"(b.z),"f"(b.w) : "memory");
		}
	}
}

// Variant: the gather goes global -> shared memory with cp.async (LDGSTS, 2 x 16 bytes per record, L1 bypassed, L2::64B
// hint optional), so a thread can have PT records in flight without holding registers for them; each thread then
// reads its own records back and stores them 4x.
template <int PT, bool HINT>
__global__ void __launch_bounds__(256) k_gather_cpasync(const float4 *state, const uint32_t *idx, float4 *out, uint32_t n)
{
	extern __shared__ float4 sm[];
	const uint32_t base = blockIdx.x * 256 * PT;
#pragma unroll
	for (int u = 0; u < PT; u++) {
		uint32_t i = base + u * 256 + threadIdx.x;
		if (i < n) {
			const float4 *src = state + 2ull * idx[i];
			uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm + 2 * (u * 256 + threadIdx.x));
			if (HINT) {
				asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
				asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" :: "r"(dst + 16), "l"(src + 1) : "memory");
			} else {
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + 16), "l"(src + 1) : "memory");
			}
		}
	}
	asm volatile("cp.async.commit_group;" ::: "memory");
	asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
	for (int u = 0; u < PT; u++) {
		uint32_t i = base + u * 256 + threadIdx.x;
		if (i >= n) continue;
		float4 a = sm[2 * (u * 256 + threadIdx.x)], b = sm[2 * (u * 256 + threadIdx.x) + 1];
		float4 *d = out + 8ull * i;
		asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(d), "f"(a.x),"f"(a.y),"f"(a.z),"f"(a.w),"f"(b.x),"f"(b.y),"f"(b.z),"f"(b.w) : "memory");
		asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(d + 2), "f"(a.x),"f"(a.y),"f"(a.z),"f"(a.w),"f"(b.x),"f"(b.y),"f"(b.z),"f"(b.w) : "memory");
		asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(d + 4), "f"(a.x),"f"(a.y),"f"(a.z),"f"(a.w),"f"(b.x),"f"(b.y),"f"(b.z),"f"(b.w) : "memory");
		asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(d + 6), "f"(a.x),"f"(a.y),"f"(a.z),"f"(a.w),"f"(b.x),"f"(b.y),"f"(b.z),"f"(b.w) : "memory");
	}
}

template <int PT, bool HINT>
float run_cpasync(const float4 *state, const uint32_t *idx, float4 *out, uint32_t n)
{
	const size_t smem = (size_t)256 * PT * 32;
	cudaFuncSetAttribute(k_gather_cpasync<PT, HINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	uint32_t grid = (n + 256 * PT - 1) / (256 * PT);
	for (int i = 0; i < 2; i++) k_gather_cpasync<PT, HINT><<<grid, 256, smem>>>(state, idx, out, n);
	cudaEventRecord(e0);
	for (int i = 0; i < 5; i++) k_gather_cpasync<PT, HINT><<<grid, 256, smem>>>(state, idx, out, n);
	cudaEventRecord(e1); CK(cudaDeviceSynchronize());
	float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}

template <int MODE, bool W4>
float run(const float4 *state, const uint32_t *idx, float4 *out, uint32_t n)
{
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int i = 0; i < 2; i++) k_gather<MODE, W4><<<(n + 1023) / 1024, 256>>>(state, idx, out, n);
	cudaEventRecord(e0);
	for (int i = 0; i < 5; i++) k_gather<MODE, W4><<<(n + 1023) / 1024, 256>>>(state, idx, out, n);
	cudaEventRecord(e1); CK(cudaDeviceSynchronize());
	float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}

int main()
{
	const uint32_t n = 10000000;
	std::vector<uint32_t> h(n);
	for (uint32_t i = 0; i < n; i++) h[i] = i;
	std::mt19937 rng(1); std::shuffle(h.begin(), h.end(), rng);
	float4 *state, *out; uint32_t *idx;
	CK(cudaMalloc(&state, (size_t)n * 32)); CK(cudaMalloc(&out, (size_t)n * 128)); CK(cudaMalloc(&idx, (size_t)n * 4));
	CK(cudaMemset(state, 0, (size_t)n * 32));
	CK(cudaMemcpy(idx, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
	size_t lim = 0; cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity); printf("default L2 fetch granularity limit: %zu\n", lim);
	for (size_t gran : { (size_t)0, (size_t)32, (size_t)64, (size_t)128 }) {
		if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity); printf("set %zu -> %s, now %zu\n", gran, cudaGetErrorString(e), lim); }
		printf("  gather+1x store : v8 %.3f  cg %.3f  cs %.3f  ldg2x %.3f  nc.noalloc %.3f  lu %.3f ms\n",
			run<0, false>(state, idx, out, n), run<1, false>(state, idx, out, n), run<2, false>(state, idx, out, n), run<3, false>(state, idx, out, n), run<4, false>(state, idx, out, n), run<5, false>(state, idx, out, n));
		printf("  gather+4x store : v8 %.3f  cg %.3f  cs %.3f  ldg2x %.3f  nc.noalloc %.3f  lu %.3f ms\n",
			run<0, true>(state, idx, out, n), run<1, true>(state, idx, out, n), run<2, true>(state, idx, out, n), run<3, true>(state, idx, out, n), run<4, true>(state, idx, out, n), run<5, true>(state, idx, out, n));
	}
	printf("  gather+4x store, hints: L2::64B v8 %.3f  L2::64B 2xv4 %.3f  no_allocate+evict_first %.3f  volatile 2xv4 %.3f ms\n",
		run<6, true>(state, idx, out, n), run<7, true>(state, idx, out, n), run<8, true>(state, idx, out, n), run<9, true>(state, idx, out, n));
	printf("  gather+4x store through cp.async -> smem: PT=4 %.3f (L2::64B %.3f)  PT=8 %.3f (L2::64B %.3f)  PT=16 %.3f (L2::64B %.3f) ms\n",
		run_cpasync<4, false>(state, idx, out, n), run_cpasync<4, true>(state, idx, out, n), run_cpasync<8, false>(state, idx, out, n), run_cpasync<8, true>(state, idx, out, n),
		run_cpasync<16, false>(state, idx, out, n), run_cpasync<16, true>(state, idx, out, n));
	// sorted (identity) order for reference
	for (uint32_t i = 0; i < n; i++) h[i] = i;
	CK(cudaMemcpy(idx, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
	printf("  identity order  : 1x %.3f  4x %.3f ms\n", run<1, false>(state, idx, out, n), run<1, true>(state, idx, out, n));
	return 0;
}
