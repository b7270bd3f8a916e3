// Micro-benchmark: sorted expansion as C passes over L2-sized slot chunks (gathers hit L2) vs one pass of random gathers.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ inline void ld256cg(const float4 *p, float4 &a, float4 &b) { asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p)); }
__device__ inline void ld256el(const float4 *p, float4 &a, float4 &b, unsigned long long pol) { asm volatile("ld.global.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p), "l"(pol)); }
__device__ inline void st256cs(float4 *p, float4 a, float4 b) { asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(a.x),"f"(a.y),"f"(a.z),"f"(a.w),"f"(b.x),"f"(b.y),"f"(b.z),"f"(b.w) : "memory"); }

template <bool HINT>
__global__ void __launch_bounds__(256) k_chunk(const float4 *state, const uint32_t *vals, float4 *out, uint32_t n, uint32_t lo, uint32_t hi)
{
	unsigned long long pol = 0;
	if (HINT) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	uint32_t i = blockIdx.x * 1024 + threadIdx.x;
	uint32_t slot[4]; float4 a[4], b[4];
#pragma unroll
	for (int u = 0; u < 4; u++) { uint32_t j = i + u * 256; slot[u] = j < n ? vals[j] : 0xffffffffu; if (slot[u] < lo || slot[u] >= hi) slot[u] = 0xffffffffu; }
#pragma unroll
	for (int u = 0; u < 4; u++) if (slot[u] != 0xffffffffu) { if (HINT) ld256el(state + 2ull * slot[u], a[u], b[u], pol); else ld256cg(state + 2ull * slot[u], a[u], b[u]); }
#pragma unroll
	for (int u = 0; u < 4; u++) {
		if (slot[u] == 0xffffffffu) continue;
		float4 *d = out + 8ull * (i + u * 256);
		st256cs(d, a[u], b[u]); st256cs(d + 2, a[u], b[u]); st256cs(d + 4, a[u], b[u]); st256cs(d + 6, a[u], b[u]);
	}
}

int main()
{
	const uint32_t n = 10000000;
	std::vector<uint32_t> h(n);
	for (uint32_t i = 0; i < n; i++) h[i] = i;
	std::mt19937 rng(1); std::shuffle(h.begin(), h.end(), rng);
	float4 *state, *out; uint32_t *vals;
	CK(cudaMalloc(&state, (size_t)n * 32)); CK(cudaMalloc(&out, (size_t)n * 128)); CK(cudaMalloc(&vals, (size_t)n * 4));
	CK(cudaMemset(state, 0, (size_t)n * 32));
	CK(cudaMemcpy(vals, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int hint = 0; hint < 2; hint++) for (int chunks : { 1, 2, 4, 6, 8 }) {
		float best = 1e9f;
		for (int rep = 0; rep < 4; rep++) {
			cudaEventRecord(e0);
			for (int c = 0; c < chunks; c++) {
				uint32_t lo = (uint32_t)((uint64_t)n * c / chunks), hi = (uint32_t)((uint64_t)n * (c + 1) / chunks);
				if (hint) k_chunk<true><<<(n + 1023) / 1024, 256>>>(state, vals, out, n, lo, hi);
				else k_chunk<false><<<(n + 1023) / 1024, 256>>>(state, vals, out, n, lo, hi);
			}
			cudaEventRecord(e1); CK(cudaDeviceSynchronize());
			float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
		}
		printf("hint=%d chunks=%d : %.3f ms\n", hint, chunks, best);
	}
	return 0;
}
