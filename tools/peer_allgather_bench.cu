// Micro-benchmark (N GPUs, one process): every GPU pulls the run of every GPU (its own included) out of peer memory
// and writes it 4x expanded into its own draw stream -- sprExpandFromPeers on all ranks at once. What limits the
// ingress rate when all N GPUs pull from all N at the same time?
//   order 0  interleaved (product): work item j = chunk j / N of rank (j + me) % N
//   order 1  phased: rank (me + k) % N for k = 0..N-1, one source at a time (at any moment every source has one reader)
//   order 2  phased, two sources at a time (k and k + N/2)
//   stores 0 read-only (xor-reduce): the NVLink side alone;  stores 1 with the 4x expansion into local HBM
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peer_allgather_bench tools/peer_allgather_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Runs { const float4 *buf[16]; uint32_t n, me; };

__device__ inline void st_rec(float4 *p, const float4 &a, const float4 &b)
{
	asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(a.x),"f"(a.y),"f"(a.z),"f"(a.w),"f"(b.x),"f"(b.y),"f"(b.z),"f"(b.w) : "memory");
}
__device__ inline void ld_rec(const float4 *p, float4 &a, float4 &b)
{
	asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p));
}

template <bool STORES>
__device__ inline void do_chunk(const float4 *src, float4 *dst, uint64_t base, uint64_t count, float &acc)
{
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	float4 a[4], b[4];
#pragma unroll
	for (int u = 0; u < 4; u++) {
		a[u] = make_float4(0, 0, 0, 0); b[u] = a[u];
		uint64_t i = base + u * 256 + threadIdx.x;
		if (i < count) ld_rec(src + 2 * i, a[u], b[u]);
	}
	if (!STORES) {
#pragma unroll
		for (int u = 0; u < 4; u++) acc += a[u].x + b[u].w;
		return;
	}
#pragma unroll
	for (int u = 0; u < 4; u++) {
		const uint64_t wbase = base + u * 256 + warp * 32;
		if (wbase >= count) continue;
		const uint32_t cnt = count - wbase < 32 ? (uint32_t)(count - wbase) : 32u;
		float4 *d = dst + 8ull * wbase;
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const uint32_t p = k * 8 + (lane >> 2);
			float4 x, y;
			x.x = __shfl_sync(0xffffffffu, a[u].x, p); x.y = __shfl_sync(0xffffffffu, a[u].y, p); x.z = __shfl_sync(0xffffffffu, a[u].z, p); x.w = __shfl_sync(0xffffffffu, a[u].w, p);
			y.x = __shfl_sync(0xffffffffu, b[u].x, p); y.y = __shfl_sync(0xffffffffu, b[u].y, p); y.z = __shfl_sync(0xffffffffu, b[u].z, p); y.w = __shfl_sync(0xffffffffu, b[u].w, p);
			if (p < cnt) st_rec(d + 2 * (k * 32 + lane), x, y);
		}
	}
}

template <bool STORES>
__global__ void __launch_bounds__(256) k_pull(Runs runs, float4 *out, uint64_t count, int order, float *sink)
{
	const uint32_t N = runs.n;
	const uint64_t chunks = (count + 1023) / 1024;
	float acc = 0.0f;
	if (order == 0) {
		for (uint64_t j = blockIdx.x; j < chunks * N; j += gridDim.x) {
			const uint32_t r = (uint32_t)((j + runs.me) % N);
			do_chunk<STORES>(runs.buf[r], out + 8ull * r * count, (j / N) * 1024, count, acc);
		}
	} else {
		const uint32_t width = order == 1 ? 1 : 2;
		for (uint32_t k = 0; k < N; k += width) {
			for (uint64_t j = blockIdx.x; j < chunks * width; j += gridDim.x) {
				const uint32_t which = (uint32_t)(j % width);
				const uint32_t r = (runs.me + (which ? k / 2 + N / 2 : (width == 2 ? k / 2 : k))) % N;
				do_chunk<STORES>(runs.buf[r], out + 8ull * r * count, (j / width) * 1024, count, acc);
			}
		}
	}
	if (acc == 123.456f) *sink = acc;
}

int main(int argc, char **argv)
{
	int ndev = 0; CK(cudaGetDeviceCount(&ndev));
	if (ndev > 16) ndev = 16;
	const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 8388608ull; // one C4 slice per GPU: 268 MB of records
	printf("%d GPUs, %llu records (%.0f MB) per GPU; every GPU pulls all %d runs: %.2f GB over NVLink into each GPU, %.2f GB written locally\n",
		ndev, (unsigned long long)n, n * 32 / 1e6, ndev, (ndev - 1) * n * 32 / 1e9, ndev * n * 128 / 1e9);
	std::vector<float4*> src(ndev), out(ndev); std::vector<float*> sink(ndev); std::vector<cudaStream_t> st(ndev);
	std::vector<cudaEvent_t> e0(ndev), e1(ndev);
	for (int d = 0; d < ndev; d++) {
		CK(cudaSetDevice(d));
		for (int p = 0; p < ndev; p++) if (p != d) { int can = 0; CK(cudaDeviceCanAccessPeer(&can, d, p)); if (!can) { printf("no peer access %d -> %d\n", d, p); return 1; } CK(cudaDeviceEnablePeerAccess(p, 0)); }
		CK(cudaMalloc(&src[d], n * 32)); CK(cudaMemset(src[d], d + 1, n * 32));
		CK(cudaMalloc(&out[d], (uint64_t)ndev * n * 128)); CK(cudaMalloc(&sink[d], 4));
		CK(cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking));
		CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d]));
	}
	cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
	const int sms = prop.multiProcessorCount;
	const int reps = 5;
	for (int stores = 0; stores < 2; stores++) for (int order = 0; order < 3; order++) for (int per : { 2, 8 }) {
		if (ndev < 4 && order == 2) continue;
		for (int pass = 0; pass < 2; pass++) { // pass 0 = warm-up
			for (int d = 0; d < ndev; d++) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
			for (int d = 0; d < ndev; d++) { CK(cudaSetDevice(d)); CK(cudaEventRecord(e0[d], st[d])); }
			for (int i = 0; i < (pass ? reps : 1); i++) for (int d = 0; d < ndev; d++) {
				CK(cudaSetDevice(d));
				Runs r; r.n = ndev; r.me = d; for (int p = 0; p < 16; p++) r.buf[p] = p < ndev ? src[p] : nullptr;
				if (stores) k_pull<true><<<sms * per, 256, 0, st[d]>>>(r, out[d], n, order, sink[d]);
				else k_pull<false><<<sms * per, 256, 0, st[d]>>>(r, out[d], n, order, sink[d]);
			}
			for (int d = 0; d < ndev; d++) { CK(cudaSetDevice(d)); CK(cudaEventRecord(e1[d], st[d])); }
			float worst = 0;
			for (int d = 0; d < ndev; d++) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); CK(cudaGetLastError()); float ms; CK(cudaEventElapsedTime(&ms, e0[d], e1[d])); worst = std::max(worst, ms); }
			if (pass) {
				const float ms = worst / reps;
				printf("  %s, order %d (%s), %d CTAs/SM: %.3f ms per all-gather, %.0f GB/s into each GPU\n", stores ? "pull + 4x expand" : "pull only      ", order,
					order == 0 ? "interleaved" : order == 1 ? "one source at a time" : "two sources at a time", per, ms, (ndev - 1) * n * 32 / (ms * 1e-3) / 1e9);
			}
		}
	}
	return 0;
}
