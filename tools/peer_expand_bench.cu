// Micro-benchmark (2 GPUs, one process): how fast can GPU 0 pull a run of 32-byte records out of GPU 1's memory over
// NVLink and write it 4x expanded into its own memory -- the inner loop of sprExpandFromPeers.
//   copy      cudaMemcpyPeerAsync of the run (copy engine): the link's practical ceiling
//   read      kernel that only reads the peer run (256-bit loads, xor-reduced): NVLink read rate of SM loads
//   lsu       the product pattern: 4 x ld.v8 in flight per thread, shuffle transpose, st.v8 (k_expand_from_peers)
//   lsu-lean  the same with fewer resident threads (CTAs per SM as given)
//   tma       cp.async.bulk (peer -> shared memory, mbarrier) + 4 x cp.async.bulk.tensor.3d stores per 128 records,
//             one thread per pipeline, STAGES x 4 KB in flight per warp: no LSU instruction touches the data
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peer_expand_bench tools/peer_expand_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ inline uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline void st_rec(float4 *p, const float4 &a, const float4 &b)
{
	asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(a.x),"f"(a.y),"f"(a.z),"f"(a.w),"f"(b.x),"f"(b.y),"f"(b.z),"f"(b.w) : "memory");
}
__device__ inline void ld_rec(const float4 *p, float4 &a, float4 &b)
{
	asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x),"=f"(a.y),"=f"(a.z),"=f"(a.w),"=f"(b.x),"=f"(b.y),"=f"(b.z),"=f"(b.w) : "l"(p));
}

__global__ void __launch_bounds__(256) k_read(const float4 *src, uint64_t n, float *sink)
{
	float acc = 0.0f;
	for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
		float4 a, b; ld_rec(src + 2 * i, a, b);
		acc += a.x + b.w;
	}
	if (acc == 123.456f) *sink = acc;
}

__global__ void __launch_bounds__(256) k_lsu(const float4 *src, uint64_t n, float4 *out)
{
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	for (uint64_t base = (uint64_t)blockIdx.x * 1024; base < n; base += (uint64_t)gridDim.x * 1024) {
		float4 a[4], b[4];
#pragma unroll
		for (int u = 0; u < 4; u++) {
			a[u] = make_float4(0, 0, 0, 0); b[u] = a[u];
			uint64_t i = base + u * 256 + threadIdx.x;
			if (i < n) ld_rec(src + 2 * i, a[u], b[u]);
		}
#pragma unroll
		for (int u = 0; u < 4; u++) {
			const uint64_t wbase = base + u * 256 + warp * 32;
			if (wbase >= n) continue;
			const uint32_t cnt = n - wbase < 32 ? (uint32_t)(n - wbase) : 32u;
			float4 *d = out + 8ull * wbase;
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
}

// One pipeline per warp (lane 0 drives it): STAGES chunks of 128 records (4 KB) in flight.
template <int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32) k_tma(const float4 *src, uint64_t n, float4 *out, const __grid_constant__ CUtensorMap map)
{
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ __align__(8) unsigned long long bars[WARPS][STAGES];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	if (lane != 0) return;
	uint8_t *stage = smem + (size_t)warp * STAGES * 4096;
	for (int s = 0; s < STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[warp][s])) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	const uint64_t chunks = n / 128; // the tail (< 128 records) is left to the caller in this benchmark
	const uint64_t stride = (uint64_t)gridDim.x * WARPS;
	const uint64_t first = (uint64_t)blockIdx.x * WARPS + warp;
	// prologue: STAGES loads in flight
	uint64_t issue = first;
	for (int s = 0; s < STAGES; s++, issue += stride) {
		if (issue >= chunks) break;
		const uint32_t bar = smem_u32(&bars[warp][s]);
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(4096) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			:: "r"(smem_u32(stage + s * 4096)), "l"(src + 256ull * issue), "r"(4096), "r"(bar) : "memory");
	}
	uint32_t phase = 0;
	int s = 0;
	for (uint64_t c = first; c < chunks; c += stride) {
		const uint32_t bar = smem_u32(&bars[warp][s]);
		uint32_t done = 0;
		while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
#pragma unroll
		for (int k = 0; k < 4; k++)
			asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
				:: "l"((uint64_t)&map), "r"(smem_u32(stage + s * 4096)), "r"(0), "r"(k), "r"((int)(c * 128)) : "memory");
		asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		// refill this stage once its stores have read it
		if (issue < chunks) {
			asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(4096) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				:: "r"(smem_u32(stage + s * 4096)), "l"(src + 256ull * issue), "r"(4096), "r"(bar) : "memory");
			issue += stride;
		}
		if (++s == STAGES) { s = 0; phase ^= 1u; }
	}
	asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
	CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make_map(float4 *out, uint64_t n)
{
	void *fn = nullptr; cudaDriverEntryPointQueryResult qr;
	CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
	CUtensorMap m;
	cuuint64_t dims[3] = { 8, 4, n }; cuuint64_t strides[2] = { 32, 128 }; cuuint32_t box[3] = { 8, 1, 128 }; cuuint32_t es[3] = { 1, 1, 1 };
	CUresult r = ((EncodeFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
		CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
	return m;
}

template <typename F>
static float timeit(F f, int reps = 10)
{
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int i = 0; i < 2; i++) f();
	CK(cudaDeviceSynchronize());
	cudaEventRecord(e0);
	for (int i = 0; i < reps; i++) f();
	cudaEventRecord(e1); CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
	float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / reps;
}

static size_t count_bad(const float4 *out, uint64_t n)
{
	std::vector<float> o((size_t)n * 32);
	CK(cudaMemcpy(o.data(), out, (size_t)n * 128, cudaMemcpyDeviceToHost));
	size_t bad = 0;
	for (uint64_t i = 0; i < n; i++) for (int c = 0; c < 4; c++) if (o[i * 32 + c * 8] != (float)(i & 0xffffff) || o[i * 32 + c * 8 + 7] != (float)(i & 1023)) bad++;
	return bad;
}

int main(int argc, char **argv)
{
	int ndev = 0; CK(cudaGetDeviceCount(&ndev));
	const int srcDev = ndev > 1 ? 1 : 0;
	printf("%d device(s); records live on device %d, kernels run on device 0%s\n", ndev, srcDev, ndev > 1 ? "" : " (LOCAL run: no NVLink in this measurement)");
	const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 8388608ull; // one C4 slice: 128 x 65536 records = 268 MB
	float4 *src, *out, *stagebuf; float *sink;
	CK(cudaSetDevice(srcDev));
	CK(cudaMalloc(&src, n * 32));
	{
		std::vector<float> s((size_t)n * 8, 0.0f);
		for (uint64_t i = 0; i < n; i++) { s[i * 8] = (float)(i & 0xffffff); s[i * 8 + 7] = (float)(i & 1023); }
		CK(cudaMemcpy(src, s.data(), n * 32, cudaMemcpyHostToDevice));
	}
	CK(cudaSetDevice(0));
	if (ndev > 1) { int can = 0; CK(cudaDeviceCanAccessPeer(&can, 0, 1)); if (!can) { printf("no peer access\n"); return 1; } CK(cudaDeviceEnablePeerAccess(1, 0)); }
	CK(cudaMalloc(&out, n * 128)); CK(cudaMalloc(&stagebuf, n * 32)); CK(cudaMalloc(&sink, 4));
	cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
	const int sms = prop.multiProcessorCount;
	CUtensorMap map = make_map(out, n);
	auto gbs = [&](float ms) { return (double)n * 32 / (ms * 1e-3) / 1e9; };

	float ms = timeit([&] { CK(cudaMemcpyPeerAsync(stagebuf, 0, src, srcDev, n * 32, 0)); });
	printf("  copy engine (cudaMemcpyPeerAsync)        : %.4f ms  %.0f GB/s in\n", ms, gbs(ms));
	for (int per : { 2, 4, 8 }) {
		ms = timeit([&] { k_read<<<sms * per, 256>>>(src, n, sink); });
		printf("  read-only kernel, %d CTAs/SM               : %.4f ms  %.0f GB/s in\n", per, ms, gbs(ms));
	}
	for (int per : { 1, 2, 4, 8 }) {
		ms = timeit([&] { k_lsu<<<sms * per, 256>>>(src, n, out); });
		printf("  lsu expand, %d CTAs/SM x 256 threads       : %.4f ms  %.0f GB/s in\n", per, ms, gbs(ms));
	}
	CK(cudaMemset(out, 0xff, n * 128));
	k_lsu<<<sms * 8, 256>>>(src, n, out); CK(cudaDeviceSynchronize());
	printf("  check lsu: %zu bad copies\n", count_bad(out, n));
#define RUN_TMA(W, S, PER) do { \
		CK(cudaFuncSetAttribute(k_tma<W, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, W * S * 4096)); \
		ms = timeit([&] { k_tma<W, S><<<sms * PER, W * 32, W * S * 4096>>>(src, n, out, map); }); \
		printf("  tma expand, %d warps x %d stages x %d CTAs/SM : %.4f ms  %.0f GB/s in  (%d KB in flight per SM)\n", W, S, PER, ms, gbs(ms), W * S * PER * 4); } while (0)
	RUN_TMA(4, 2, 1); RUN_TMA(4, 4, 1); RUN_TMA(8, 4, 1); RUN_TMA(8, 6, 1); RUN_TMA(4, 4, 2); RUN_TMA(4, 4, 3); RUN_TMA(2, 4, 4); RUN_TMA(1, 4, 8);
	CK(cudaMemset(out, 0xff, n * 128));
	k_tma<4, 4><<<sms, 128, 4 * 4 * 4096>>>(src, n, out, map); CK(cudaDeviceSynchronize());
	printf("  check tma: %zu bad copies\n", count_bad(out, n / 128 * 128));
	return 0;
}
