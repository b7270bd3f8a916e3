// Micro-benchmark: the sorted expansion's STORE pattern. One lane per particle holds a 32-byte record and writes it
// four times (128 contiguous bytes per particle). Variant A is the product pattern (4 x st.v8 per lane: each store
// instruction of a warp touches 32 different 128-byte lines); B stages the warp's 32 records in shared memory and
// writes them back transposed (each store instruction of a warp covers 1 KB contiguous); C does the same with shuffles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/expand_store_bench tools/expand_store_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ inline void ld_rec(const float4 *p, float4 &a, float4 &b, bool hint64)
{
	if (hint64) asm volatile("ld.global.L2::64B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p));
	else asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p));
}
__device__ inline void st_rec(float4 *p, const float4 &a, const float4 &b)
{
	asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(a.x),"f"(a.y),"f"(a.z),"f"(a.w),"f"(b.x),"f"(b.y),"f"(b.z),"f"(b.w) : "memory");
}

// VARIANT 0: strided (product), 1: shared-memory transpose, 2: shuffle transpose
template <int VARIANT, bool HINT>
__global__ void __launch_bounds__(256) k_expand(const float4 *state, const uint32_t *idx, float4 *out, uint32_t n)
{
	__shared__ float4 stage[8][64]; // one 1 KB buffer per warp
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t base = blockIdx.x * 1024;
	uint32_t slot[4];
	float4 a[4], b[4];
#pragma unroll
	for (int u = 0; u < 4; u++) { uint32_t i = base + u * 256 + threadIdx.x; slot[u] = i < n ? idx[i] : 0xffffffffu; }
#pragma unroll
	for (int u = 0; u < 4; u++) if (slot[u] != 0xffffffffu) ld_rec(state + 2ull * slot[u], a[u], b[u], HINT);
#pragma unroll
	for (int u = 0; u < 4; u++) {
		const uint32_t wbase = base + u * 256 + warp * 32; // first particle of this warp's group
		if (wbase >= n) continue;
		const uint32_t cnt = min(32u, n - wbase);
		float4 *d = out + 8ull * wbase;
		if (VARIANT == 0) {
			if (lane < cnt) { float4 *p = d + 8 * lane; st_rec(p, a[u], b[u]); st_rec(p + 2, a[u], b[u]); st_rec(p + 4, a[u], b[u]); st_rec(p + 6, a[u], b[u]); }
		} else if (VARIANT == 1) {
			stage[warp][2 * lane] = a[u]; stage[warp][2 * lane + 1] = b[u];
			__syncwarp();
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const uint32_t p = j * 8 + (lane >> 2); // particle of the group whose copy (lane & 3) this lane writes
				float4 x = stage[warp][2 * p], y = stage[warp][2 * p + 1];
				if (p < cnt) st_rec(d + 2 * (j * 32 + lane), x, y);
			}
			__syncwarp();
		} else {
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
}

template <int VARIANT, bool HINT>
float run(const float4 *state, const uint32_t *idx, float4 *out, uint32_t n)
{
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int i = 0; i < 2; i++) k_expand<VARIANT, HINT><<<(n + 1023) / 1024, 256>>>(state, idx, out, n);
	cudaEventRecord(e0);
	for (int i = 0; i < 10; i++) k_expand<VARIANT, HINT><<<(n + 1023) / 1024, 256>>>(state, idx, out, n);
	cudaEventRecord(e1); CK(cudaDeviceSynchronize());
	float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 10;
}

static void check(const float4 *state, const uint32_t *idx, float4 *out, uint32_t n, const std::vector<uint32_t> &h)
{
	// records hold their own index in .x: every copy of output particle i must carry h[i]
	for (int v = 0; v < 3; v++) {
		CK(cudaMemset(out, 0xff, (size_t)n * 128));
		if (v == 0) k_expand<0, false><<<(n + 1023) / 1024, 256>>>(state, idx, out, n);
		if (v == 1) k_expand<1, false><<<(n + 1023) / 1024, 256>>>(state, idx, out, n);
		if (v == 2) k_expand<2, false><<<(n + 1023) / 1024, 256>>>(state, idx, out, n);
		CK(cudaDeviceSynchronize());
		std::vector<float> o((size_t)n * 32);
		CK(cudaMemcpy(o.data(), out, (size_t)n * 128, cudaMemcpyDeviceToHost));
		size_t bad = 0;
		for (uint32_t i = 0; i < n; i++) for (int c = 0; c < 4; c++) { if (o[(size_t)i * 32 + c * 8] != (float)h[i] || o[(size_t)i * 32 + c * 8 + 7] != (float)(h[i] & 1023)) bad++; }
		printf("  check variant %d: %zu bad copies of %zu\n", v, bad, (size_t)n * 4);
	}
}

int main()
{
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
		CK(cudaMemcpy(idx, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
		printf("n = %u (random order)\n", n);
		if (n < 2000000u) check(state, idx, out, n - 7, h); // ragged tail on purpose
		printf("  strided stores (product)   : cg %.4f  L2::64B %.4f ms\n", run<0, false>(state, idx, out, n), run<0, true>(state, idx, out, n));
		printf("  shared-memory transpose    : cg %.4f  L2::64B %.4f ms\n", run<1, false>(state, idx, out, n), run<1, true>(state, idx, out, n));
		printf("  shuffle transpose          : cg %.4f  L2::64B %.4f ms\n", run<2, false>(state, idx, out, n), run<2, true>(state, idx, out, n));
		for (uint32_t i = 0; i < n; i++) h[i] = i;
		CK(cudaMemcpy(idx, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
		printf("n = %u (identity order)\n", n);
		printf("  strided stores (product)   : cg %.4f ms\n", run<0, false>(state, idx, out, n));
		printf("  shared-memory transpose    : cg %.4f ms\n", run<1, false>(state, idx, out, n));
		printf("  shuffle transpose          : cg %.4f ms\n", run<2, false>(state, idx, out, n));
		cudaFree(state); cudaFree(out); cudaFree(idx);
	}
	return 0;
}
