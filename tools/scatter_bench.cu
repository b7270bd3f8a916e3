// Micro-benchmark: expansion by SLOT order with scattered 128-byte line writes (alternative to the rank-order gather).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ inline void ld256(const float4 *p, float4 &a, float4 &b) { asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p)); }
__device__ inline void st256(float4 *p, float4 a, float4 b) { asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(a.x),"f"(a.y),"f"(a.z),"f"(a.w),"f"(b.x),"f"(b.y),"f"(b.z),"f"(b.w) : "memory"); }

// sequential read of record i, scattered write of 4 copies to out[perm[i]]
__global__ void __launch_bounds__(256) k_scatter4(const float4 *state, const uint32_t *perm, float4 *out, uint32_t n)
{
	uint32_t i = blockIdx.x * 1024 + threadIdx.x;
#pragma unroll
	for (int u = 0; u < 4; u++, i += 256) {
		if (i >= n) return;
		float4 a, b; ld256(state + 2ull * i, a, b);
		float4 *d = out + 8ull * perm[i];
		st256(d, a, b); st256(d + 2, a, b); st256(d + 4, a, b); st256(d + 6, a, b);
	}
}
// 8 lanes per particle cooperate on the 128-byte line (one 16-byte store each)
__global__ void __launch_bounds__(256) k_scatter4_coop(const float4 *state, const uint32_t *perm, float4 *out, uint32_t n)
{
	uint64_t q = (uint64_t)blockIdx.x * 256 + threadIdx.x;  // float4 index of the output
	for (int u = 0; u < 8; u++, q += (uint64_t)gridDim.x * 256) {
		uint64_t i = q >> 3;
		if (i >= n) return;
		float4 v = state[2 * i + (q & 1)];
		out[8ull * perm[i] + (q & 7)] = v;
	}
}
// random 4-byte scatter: rankOf[slot[r]] = r
__global__ void k_inv(const uint32_t *perm, uint32_t *inv, uint32_t n)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) inv[perm[i]] = i;
}

template <typename F> float timeit(F f) {
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	f(); f();
	cudaEventRecord(e0); for (int i = 0; i < 5; i++) f(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
	float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}

int main()
{
	const uint32_t n = 10000000;
	std::vector<uint32_t> h(n);
	for (uint32_t i = 0; i < n; i++) h[i] = i;
	std::mt19937 rng(1); std::shuffle(h.begin(), h.end(), rng);
	float4 *state, *out; uint32_t *perm, *inv;
	CK(cudaMalloc(&state, (size_t)n * 32)); CK(cudaMalloc(&out, (size_t)n * 128)); CK(cudaMalloc(&perm, (size_t)n * 4)); CK(cudaMalloc(&inv, (size_t)n * 4));
	CK(cudaMemset(state, 0, (size_t)n * 32));
	CK(cudaMemcpy(perm, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
	printf("scatter 4x (lane per particle, 4 x 256-bit stores) : %.3f ms\n", timeit([&] { k_scatter4<<<(n + 1023) / 1024, 256>>>(state, perm, out, n); }));
	printf("scatter 4x (8 lanes per particle, 128-bit stores)   : %.3f ms\n", timeit([&] { k_scatter4_coop<<<(uint32_t)(((uint64_t)n * 8 + 2047) / 2048), 256>>>(state, perm, out, n); }));
	printf("inverse permutation (random 4-byte scatter)          : %.3f ms\n", timeit([&] { k_inv<<<(n + 255) / 256, 256>>>(perm, inv, n); }));
	return 0;
}
