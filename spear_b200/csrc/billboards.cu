// billboards.cu -- SURVEY.md 8(f) row N4: cl::BillboardSystem::renderMain
// (src/client/BillboardSystem.cpp:119-228) on the GPU: stable sort of the frame's billboards by
// (depth, insertion index) (BillboardOrder, :41-51), cut to the frame budget (:121-123), crop
// against the sprite quad, skip empty crops (:136-138), and expand every remaining billboard to
// 4 x {position3, texCoord2, rgba8} = 96 bytes in sorted order (:140-196). The atlas of every
// emitted quad is returned too; the host turns runs of equal atlases into draws (:198-203).
//
// Unlike the particle path this one is stateless: billboards are re-added every frame
// (addBillboard, :80-117), so a frame is upload -> keys -> 4 radix passes -> flags -> scan ->
// expand. The sort is a plain segmented-by-nothing LSD radix sort (per-tile digit counts, one scan
// over tiles x digits, stable scatter); a few thousand billboards do not need a chained scan.

#include <cuda_runtime.h>
#include <stdint.h>

#include "device_types.h"
#include "kernels.h"

namespace spr {

#define FULL_MASK 0xffffffffu

static const uint32_t kBbTile = 1024;    // keys per CTA in the radix passes
static const uint32_t kBbThreads = 256;

__device__ inline uint32_t bb_enc_float(float f)
{
	if (f == 0.0f) f = 0.0f; // -0 and +0 compare equal in BillboardOrder::operator< (:46): one key for both
	uint32_t u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void k_bb_keys(const BillboardDev *bb, uint32_t n, uint32_t *keys, uint32_t *vals)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	keys[i] = bb_enc_float(bb[i].depth);
	vals[i] = i;
}

// per-tile digit counts, digit-major: tileHist[digit * numTiles + tile]
__global__ void __launch_bounds__(kBbThreads) k_bb_hist(const uint32_t *keys, uint32_t n, uint32_t shift, uint32_t *tileHist, uint32_t numTiles)
{
	__shared__ uint32_t sHist[256];
	sHist[threadIdx.x] = 0;
	__syncthreads();
	uint32_t base = blockIdx.x * kBbTile;
	for (uint32_t i = threadIdx.x; i < kBbTile && base + i < n; i += kBbThreads) atomicAdd(&sHist[(keys[base + i] >> shift) & 255u], 1u);
	__syncthreads();
	tileHist[threadIdx.x * numTiles + blockIdx.x] = sHist[threadIdx.x];
}

// exclusive scan of `count` words by ONE CTA (count = 256 * numTiles, or the block counts of the compaction)
__global__ void __launch_bounds__(1024) k_bb_scan(uint32_t *data, uint32_t count, uint32_t *total)
{
	__shared__ uint32_t sWarp[32];
	__shared__ uint32_t sRun;
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) sRun = 0;
	__syncthreads();
	for (uint32_t base = 0; base < count; base += 1024) {
		uint32_t i = base + tid;
		uint32_t v = i < count ? data[i] : 0u;
		uint32_t incl = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(FULL_MASK, incl, d); if ((int)lane >= d) incl += o; }
		if (lane == 31) sWarp[warp] = incl;
		__syncthreads();
		uint32_t run = sRun;
		for (uint32_t w = 0; w < warp; w++) run += sWarp[w];
		if (i < count) data[i] = run + incl - v;
		__syncthreads();
		if (tid == 1023) sRun = run + incl;
		__syncthreads();
	}
	if (tid == 0 && total) *total = sRun;
}

// stable scatter of one tile: warp w owns keys [w * 128, w * 128 + 128) of the tile, four steps of 32
__global__ void __launch_bounds__(kBbThreads) k_bb_scatter(const uint32_t *keysIn, const uint32_t *valsIn, uint32_t *keysOut, uint32_t *valsOut,
	uint32_t n, uint32_t shift, const uint32_t *tileHist, uint32_t numTiles)
{
	__shared__ uint32_t sCnt[8][256];   // per-warp digit counts -> exclusive offsets across the warps of the tile
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (uint32_t i = tid; i < 8 * 256; i += kBbThreads) (&sCnt[0][0])[i] = 0;
	__syncthreads();
	const uint32_t base = blockIdx.x * kBbTile + warp * 128;
	uint32_t key[4], rank[4];
	const uint32_t ltMask = (1u << lane) - 1u;
#pragma unroll
	for (int s = 0; s < 4; s++) {
		uint32_t i = base + s * 32 + lane;
		bool ok = i < n;
		key[s] = ok ? keysIn[i] : 0xffffffffu;
		uint32_t d = (key[s] >> shift) & 255u;
		uint32_t peers = __match_any_sync(FULL_MASK, ok ? d : 0x100u + lane);
		uint32_t before = ok ? sCnt[warp][d] : 0u;
		__syncwarp();
		rank[s] = before + __popc(peers & ltMask);
		if (ok && (peers & ltMask) == 0) sCnt[warp][d] = before + __popc(peers);
		__syncwarp();
	}
	__syncthreads();
	{
		// thread d: exclusive offsets of digit d over the warps, on top of the tile's global base for d
		uint32_t run = tileHist[tid * numTiles + blockIdx.x];
#pragma unroll
		for (int w = 0; w < 8; w++) { uint32_t c = sCnt[w][tid]; sCnt[w][tid] = run; run += c; }
	}
	__syncthreads();
#pragma unroll
	for (int s = 0; s < 4; s++) {
		uint32_t i = base + s * 32 + lane;
		if (i >= n) continue;
		uint32_t d = (key[s] >> shift) & 255u;
		uint32_t dst = sCnt[warp][d] + rank[s];
		keysOut[dst] = key[s];
		valsOut[dst] = valsIn[i];
	}
}

// :131-138 -- the cropped quad of a billboard; false when the crop is empty
__device__ __forceinline__ bool bb_crop(const BillboardDev &b, float &minX, float &minY, float &maxX, float &maxY)
{
	// sf::max / sf::min on floats are _mm_max_ss / _mm_min_ss in the reference build (src/sf/Base.h:107-108):
	// max(a, b) = a > b ? a : b, min(a, b) = a < b ? a : b (the second operand on ties and NaNs)
	minX = b.minVert[0] > b.cropMin[0] ? b.minVert[0] : b.cropMin[0]; // sf::max(sprite->minVert, billboard.cropMin), :136
	minY = b.minVert[1] > b.cropMin[1] ? b.minVert[1] : b.cropMin[1];
	maxX = b.maxVert[0] < b.cropMax[0] ? b.maxVert[0] : b.cropMax[0]; // sf::min(sprite->maxVert, billboard.cropMax), :137
	maxY = b.maxVert[1] < b.cropMax[1] ? b.maxVert[1] : b.cropMax[1];
	return !(minX >= maxX || minY >= maxY);
}

// valid flags of the first m sorted billboards -> per-block counts
__global__ void __launch_bounds__(kBbThreads) k_bb_count(const BillboardDev *bb, const uint32_t *order, uint32_t m, uint32_t *blockCounts)
{
	uint32_t i = blockIdx.x * kBbThreads + threadIdx.x;
	bool valid = false;
	if (i < m) { float a, b, c, d; valid = bb_crop(bb[order[i]], a, b, c, d); }
	uint32_t cnt = __syncthreads_count(valid);
	if (threadIdx.x == 0) blockCounts[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(kBbThreads) k_bb_expand(const BillboardDev *bb, const uint32_t *order, uint32_t m, const uint32_t *blockOffsets,
	float *verts /* 6 words per vertex */, unsigned long long *atlasOut)
{
	__shared__ uint32_t sWarp[8];
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	uint32_t i = blockIdx.x * kBbThreads + tid;
	bool valid = false;
	float minX = 0, minY = 0, maxX = 0, maxY = 0;
	uint32_t idx = 0;
	if (i < m) { idx = order[i]; valid = bb_crop(bb[idx], minX, minY, maxX, maxY); }
	uint32_t ballot = __ballot_sync(FULL_MASK, valid);
	if (lane == 0) sWarp[warp] = __popc(ballot);
	__syncthreads();
	uint32_t pos = blockOffsets[blockIdx.x] + __popc(ballot & ((1u << lane) - 1u));
	for (uint32_t w = 0; w < warp; w++) pos += sWarp[w];
	if (!valid) return;
	const BillboardDev &b = bb[idx];
	// :140-143
	float uvMinX = minX * b.vertUvScale[0] + b.vertUvBias[0];
	float uvMinY = minY * b.vertUvScale[1] + b.vertUvBias[1];
	float uvMaxX = maxX * b.vertUvScale[0] + b.vertUvBias[0];
	float uvMaxY = maxY * b.vertUvScale[1] + b.vertUvBias[1];
	// :145-146
	minX -= b.anchor[0]; minY -= b.anchor[1];
	maxX -= b.anchor[0]; maxY -= b.anchor[1];
	// sf::Mat34 is column major: t[0..2] = m00 m10 m20, t[3..5] = m01 m11 m21, t[6..8] = m02.., t[9..11] = m03 m13 m23
	const float *t = b.transform;
	const float m00 = t[0], m10 = t[1], m20 = t[2], m01 = t[3], m11 = t[4], m21 = t[5], m03 = t[9], m13 = t[10], m23 = t[11];
	// :148-161
	float xx0 = minX * m00 + m03, xx1 = maxX * m00 + m03;
	float yy0 = minY * m11 + m13, yy1 = maxY * m11 + m13;
	float xy0 = minY * m01, xy1 = maxY * m01;
	float yx0 = minX * m10, yx1 = maxX * m10;
	float zx0 = minX * m20 + m23, zx1 = maxX * m20 + m23;
	float zy0 = minY * m21, zy1 = maxY * m21;
	float *v = verts + 24ull * pos;
	const float col = __uint_as_float(b.color);
	// :165-191
	v[0] = xx0 + xy0; v[1] = yx0 + yy0; v[2] = zx0 + zy0; v[3] = uvMinX; v[4] = uvMinY; v[5] = col;
	v[6] = xx1 + xy0; v[7] = yx1 + yy0; v[8] = zx1 + zy0; v[9] = uvMaxX; v[10] = uvMinY; v[11] = col;
	v[12] = xx0 + xy1; v[13] = yx0 + yy1; v[14] = zx0 + zy1; v[15] = uvMinX; v[16] = uvMaxY; v[17] = col;
	v[18] = xx1 + xy1; v[19] = yx1 + yy1; v[20] = zx1 + zy1; v[21] = uvMaxX; v[22] = uvMaxY; v[23] = col;
	atlasOut[pos] = b.atlas;
}

// ---- the draw list (:198-203): runs of equal atlas over the expanded quads, built where the atlas ids are.
// k_bb_run_count: quad q starts a run when it is the first one or its atlas differs from quad q - 1's -> per-block counts;
// k_bb_run_write (after the scan of the counts): run r = (atlas, first quad); k_bb_run_lengths: count = next run's first
// quad - this one's. `total` (the number of quads) is only known on the device: the grids cover the frame's budget.
__global__ void __launch_bounds__(kBbThreads) k_bb_run_count(const unsigned long long *atlas, const uint32_t *total, uint32_t *blockCounts)
{
	const uint32_t q = blockIdx.x * kBbThreads + threadIdx.x, n = *total;
	const bool start = q < n && (q == 0 || atlas[q] != atlas[q - 1]);
	const uint32_t cnt = __syncthreads_count(start);
	if (threadIdx.x == 0) blockCounts[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(kBbThreads) k_bb_run_write(const unsigned long long *atlas, const uint32_t *total, const uint32_t *blockOffsets, BillboardDrawDev *draws)
{
	__shared__ uint32_t sWarp[8];
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t q = blockIdx.x * kBbThreads + tid, n = *total;
	const bool start = q < n && (q == 0 || atlas[q] != atlas[q - 1]);
	const uint32_t ballot = __ballot_sync(FULL_MASK, start);
	if (lane == 0) sWarp[warp] = __popc(ballot);
	__syncthreads();
	uint32_t pos = blockOffsets[blockIdx.x] + __popc(ballot & ((1u << lane) - 1u));
	for (uint32_t w = 0; w < warp; w++) pos += sWarp[w];
	if (start) { draws[pos].atlas = atlas[q]; draws[pos].firstQuad = q; }
}

__global__ void __launch_bounds__(256) k_bb_run_lengths(BillboardDrawDev *draws, const uint32_t *numRuns, const uint32_t *total)
{
	const uint32_t r = blockIdx.x * 256 + threadIdx.x, n = *numRuns;
	if (r >= n) return;
	draws[r].count = (r + 1 < n ? draws[r + 1].firstQuad : *total) - draws[r].firstQuad;
}

static inline uint32_t bb_div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

uint32_t launch_billboards(cudaStream_t st, const BillboardDev *bb, uint32_t n, uint32_t maxPerFrame, const BillboardScratch &sc,
	float *verts, unsigned long long *atlasOut, uint32_t *totalOut, BillboardDrawDev *drawsOut)
{
	if (!n) return 0;
	uint32_t launches = 0;
	const uint32_t numTiles = bb_div_up(n, kBbTile);
	k_bb_keys<<<bb_div_up(n, 256), 256, 0, st>>>(bb, n, sc.keys[0], sc.vals[0]); launches++;
	for (uint32_t p = 0; p < 4; p++) {
		const uint32_t a = p & 1, b = a ^ 1;
		k_bb_hist<<<numTiles, kBbThreads, 0, st>>>(sc.keys[a], n, p * 8, sc.tileHist, numTiles);
		k_bb_scan<<<1, 1024, 0, st>>>(sc.tileHist, 256 * numTiles, nullptr);
		k_bb_scatter<<<numTiles, kBbThreads, 0, st>>>(sc.keys[a], sc.vals[a], sc.keys[b], sc.vals[b], n, p * 8, sc.tileHist, numTiles);
		launches += 3;
	}
	// four passes: the sorted order is back in buffer 0
	const uint32_t m = n < maxPerFrame ? n : maxPerFrame; // :121-123
	const uint32_t blocks = bb_div_up(m, kBbThreads);
	k_bb_count<<<blocks, kBbThreads, 0, st>>>(bb, sc.vals[0], m, sc.blockCounts);
	k_bb_scan<<<1, 1024, 0, st>>>(sc.blockCounts, blocks, totalOut);
	k_bb_expand<<<blocks, kBbThreads, 0, st>>>(bb, sc.vals[0], m, sc.blockCounts, verts, atlasOut);
	// totalOut[0] = quads, totalOut[1] = draws
	k_bb_run_count<<<blocks, kBbThreads, 0, st>>>(atlasOut, totalOut, sc.blockCounts);
	k_bb_scan<<<1, 1024, 0, st>>>(sc.blockCounts, blocks, totalOut + 1);
	k_bb_run_write<<<blocks, kBbThreads, 0, st>>>(atlasOut, totalOut, sc.blockCounts, drawsOut);
	k_bb_run_lengths<<<bb_div_up(m, 256), 256, 0, st>>>(drawsOut, totalOut + 1, totalOut);
	return launches + 7;
}

} // namespace spr
