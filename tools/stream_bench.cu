// Micro-benchmark of the integrate access pattern: read 32 B/slot, write 32 B/slot in place (+4 B key).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ inline void ld256(const float4 *p, float4 &a, float4 &b) { asm("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p)); }
__device__ inline void st256(float4 *p, float4 a, float4 b) { asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(a.x),"f"(a.y),"f"(a.z),"f"(a.w),"f"(b.x),"f"(b.y),"f"(b.z),"f"(b.w) : "memory"); }

// MODE 0: 256-bit, lane stride (current kernel). 1: 2x128-bit lane stride. 2: 128-bit fully coalesced (float4 index = lane), no record semantics
template <int MODE, int R, bool INPLACE, bool KEYS>
__global__ void k_stream(float4 *state, float4 *out, uint32_t *keys, uint32_t n, float dt)
{
	uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	uint32_t base = warp * (32 * R);
	if (base >= n) return;
	float4 a[R], b[R];
	float4 *dst = INPLACE ? state : out;
	if (MODE == 2) {
		const float4 *p = state + 2ull * base;
#pragma unroll
		for (int r = 0; r < R; r++) { a[r] = p[r * 64 + lane]; b[r] = p[r * 64 + 32 + lane]; }
#pragma unroll
		for (int r = 0; r < R; r++) { a[r].x += dt; b[r].x += dt; }
		float4 *q = dst + 2ull * base;
#pragma unroll
		for (int r = 0; r < R; r++) { q[r * 64 + lane] = a[r]; q[r * 64 + 32 + lane] = b[r]; }
		return;
	}
#pragma unroll
	for (int r = 0; r < R; r++) {
		const float4 *p = state + 2ull * (base + r * 32 + lane);
		if (MODE == 0) ld256(p, a[r], b[r]); else { a[r] = p[0]; b[r] = p[1]; }
	}
#pragma unroll
	for (int r = 0; r < R; r++) {
		a[r].x += b[r].x * dt; a[r].y += b[r].y * dt; a[r].z += b[r].z * dt; a[r].w -= dt;
		float4 *p = dst + 2ull * (base + r * 32 + lane);
		if (MODE == 0) st256(p, a[r], b[r]); else { p[0] = a[r]; p[1] = b[r]; }
		if (KEYS) keys[base + r * 32 + lane] = __float_as_uint(a[r].x * a[r].x + a[r].y * a[r].y);
	}
}

template <int MODE, int R, bool INPLACE, bool KEYS>
float run(float4 *state, float4 *out, uint32_t *keys, uint32_t n, int threads)
{
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	uint32_t warps = (n + 32 * R - 1) / (32 * R);
	uint32_t blocks = (warps * 32 + threads - 1) / threads;
	for (int i = 0; i < 2; i++) k_stream<MODE, R, INPLACE, KEYS><<<blocks, threads>>>(state, out, keys, n, 0.01f);
	cudaEventRecord(e0);
	for (int i = 0; i < 10; i++) k_stream<MODE, R, INPLACE, KEYS><<<blocks, threads>>>(state, out, keys, n, 0.01f);
	cudaEventRecord(e1); CK(cudaDeviceSynchronize());
	float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 10;
}

int main()
{
	const uint32_t n = 10000000 / 128 * 128;
	float4 *state, *out; uint32_t *keys;
	CK(cudaMalloc(&state, (size_t)n * 32)); CK(cudaMalloc(&out, (size_t)n * 32)); CK(cudaMalloc(&keys, (size_t)n * 4));
	CK(cudaMemset(state, 0, (size_t)n * 32));
	printf("10M slots, 640 MB (+40 MB keys). ms (GB/s of 640 MB)\n");
#define ROW(name, call) { float ms = call; printf("  %-46s %.4f ms  %.0f GB/s\n", name, ms, 0.64 / ms * 1e3); }
	ROW("256b lane-stride R=4 inplace keys  256thr", (run<0, 4, true, true>(state, out, keys, n, 256)));
	ROW("256b lane-stride R=4 inplace       256thr", (run<0, 4, true, false>(state, out, keys, n, 256)));
	ROW("256b lane-stride R=4 out-of-place  256thr", (run<0, 4, false, false>(state, out, keys, n, 256)));
	ROW("256b lane-stride R=2 inplace       256thr", (run<0, 2, true, false>(state, out, keys, n, 256)));
	ROW("256b lane-stride R=1 inplace       256thr", (run<0, 1, true, false>(state, out, keys, n, 256)));
	ROW("256b lane-stride R=8 inplace       256thr", (run<0, 8, true, false>(state, out, keys, n, 256)));
	ROW("256b lane-stride R=4 inplace       128thr", (run<0, 4, true, false>(state, out, keys, n, 128)));
	ROW("256b lane-stride R=4 inplace       512thr", (run<0, 4, true, false>(state, out, keys, n, 512)));
	ROW("2x128b lane-stride R=4 inplace     256thr", (run<1, 4, true, false>(state, out, keys, n, 256)));
	ROW("128b coalesced     R=4 inplace     256thr", (run<2, 4, true, false>(state, out, keys, n, 256)));
	ROW("128b coalesced     R=4 out-of-place 256thr", (run<2, 4, false, false>(state, out, keys, n, 256)));
	ROW("128b coalesced     R=2 inplace     256thr", (run<2, 2, true, false>(state, out, keys, n, 256)));
	float ms; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0); for (int i = 0; i < 10; i++) cudaMemcpyAsync(out, state, (size_t)n * 32, cudaMemcpyDeviceToDevice); cudaEventRecord(e1); cudaDeviceSynchronize();
	cudaEventElapsedTime(&ms, e0, e1); printf("  %-46s %.4f ms  %.0f GB/s\n", "cudaMemcpy D2D 320 MB", ms / 10, 0.64 / (ms / 10) * 1e3);
	return 0;
}
