// Yardstick only (not product code): cub::DeviceRadixSort::SortPairs on the key shape of the 10M-particle sorted step
// (32-bit keys spanning 24 significant bits, 32-bit slot values), to know what a tuned onesweep pass costs on this GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cub_sort_bench tools/cub_sort_bench.cu
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
int main()
{
	for (uint32_t n : { 10000000u, 8388608u, 1048576u }) {
		std::vector<uint32_t> hk(n), hv(n);
		std::mt19937 rng(7);
		for (uint32_t i = 0; i < n; i++) { hk[i] = 0x41000000u + (rng() & 0xffffffu); hv[i] = i; }
		uint32_t *k0, *k1, *v0, *v1; void *tmp = nullptr; size_t tmpBytes = 0;
		CK(cudaMalloc(&k0, n * 4)); CK(cudaMalloc(&k1, n * 4)); CK(cudaMalloc(&v0, n * 4)); CK(cudaMalloc(&v1, n * 4));
		CK(cudaMemcpy(k0, hk.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(v0, hv.data(), n * 4, cudaMemcpyHostToDevice));
		for (int endBit : { 24, 32 }) {
			CK(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, k0, k1, v0, v1, (int)n, 0, endBit));
			CK(cudaMalloc(&tmp, tmpBytes));
			cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
			for (int i = 0; i < 3; i++) CK(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, k0, k1, v0, v1, (int)n, 0, endBit));
			cudaEventRecord(e0);
			for (int i = 0; i < 20; i++) CK(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, k0, k1, v0, v1, (int)n, 0, endBit));
			cudaEventRecord(e1); CK(cudaDeviceSynchronize());
			float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
			printf("n = %u, bits [0,%d): cub SortPairs %.4f ms (%d passes incl. histogram) = %.1f G pairs/s per pass\n", n, endBit, ms, endBit / 8, n / (ms * 1e-3) / 1e9 * (endBit / 8));
			CK(cudaFree(tmp)); tmp = nullptr;
		}
		cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1);
	}
	return 0;
}
