// kernels.h -- launch wrappers of kernels.cu / sort.cu (host-callable, stream-ordered).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "device_types.h"

namespace spr {

// What renderMain / the parity read-backs need from the device per effect.
struct RenderGather {
	uint32_t aliveCount, slotCount, freeCount;
	uint32_t stopEmit, firstEmit;
	float spawnTimer, emitOnTimer;
	float boundsMin[4], boundsMax[4];
};

void launch_fill_jobs(cudaStream_t st, uint32_t *dst, const FillJob *jobs, uint32_t numJobs);
void launch_scatter_words(cudaStream_t st, void *dst, const uint32_t *idx, const void *src, uint32_t n, uint32_t words);
void launch_plan(cudaStream_t st, const TablePtrs &tab, const SystemWork *work, uint32_t numWork, const ActiveEntry *active, uint32_t stamp);
void launch_emit_warp(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, const SystemWork *work, const ActiveEntry *active, uint32_t numActive, uint32_t round, uint32_t stamp);
void launch_emit_burst(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, const ActiveEntry *active, const BurstJob *jobs, uint32_t numJobs, uint32_t totalBlocks);
void launch_integrate(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, uint32_t numPoolTiles, uint32_t stamp, uint32_t round,
	uint32_t epoch, bool sortMode, bool gravityPoints, bool tinyEffects);
// ticketBase: the context's running count of tickets handed out by launches of the kernel so far (PoolPtrs::tickets)
void launch_dead_append(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, uint32_t numPoolTiles, uint32_t stamp, uint32_t round, uint32_t epoch, bool aliveScan,
	uint64_t &ticketBase);
void configure_stream_fused(int ctasPerSm[2]);
void configure_sort_kernels();
// An encoded CUtensorMap of the vertex pool, opaque to the host layer (cuda.h stays out of its headers).
struct alignas(64) VertexTensorMap { unsigned char bytes[128]; };
bool encode_vertex_map(VertexTensorMap *out, void *vertices, uint64_t slots);
// vertexMap: nullptr = expansion stores through the LSU; else full granules leave through bulk tensor stores (TMA)
uint32_t launch_stream_fused(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, uint32_t numPoolTiles, uint32_t stamp, uint32_t round,
	uint32_t epoch, bool gravityPoints, uint32_t numSms, const int ctasPerSm[2], uint64_t &ticketBase, const VertexTensorMap *vertexMap);
void launch_expand_stream(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, uint32_t numPoolTiles, uint32_t stamp);
void launch_finalize(cudaStream_t st, const TablePtrs &tab, const ActiveEntry *active, uint32_t numActive, uint32_t stamp);
void launch_query_frustum(cudaStream_t st, const TablePtrs &tab, uint32_t numGlobals, uint32_t systemIdx, const float frustum[32],
	uint32_t *outIds, uint32_t capacity, uint32_t *outCount);
void launch_gather_render(cudaStream_t st, const TablePtrs &tab, const uint32_t *effects, uint32_t n, RenderGather *out);
void launch_run_offsets(cudaStream_t st, const TablePtrs &tab, const uint32_t *effects, uint32_t n, uint64_t *offsets, uint32_t *counts);
void launch_pack_runs(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, const uint32_t *effects, uint32_t n,
	const uint64_t *offsets, float4 *out, uint64_t capacityRecords, uint32_t numSms, bool fromState);
void launch_pack_runs_buf(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, const uint32_t *effects, uint32_t n,
	const uint64_t *offsets, void *buf, uint64_t capacityRecords, uint32_t numSms, bool fromState);
void launch_expand_from_peers(cudaStream_t st, void *const *bufs, uint32_t numRanks, uint32_t myRank, float4 *out, uint64_t capacityRecords, uint64_t *totalOut, uint32_t numCtas, bool smallShare);
void launch_expand_runs(cudaStream_t st, const float4 *records, uint64_t numRecords, float4 *out, uint32_t numSms);

// vertex_stage.cu: Particle.glsl `vs` over draw records
void launch_shade_vertices(cudaStream_t st, const float4 *vertices, const ShadeJob *jobs, uint32_t numJobs, uint32_t totalBlocks,
	const uint32_t *luts, const uint8_t *fogTexels, uint32_t fogW, uint32_t fogH, float4 *out, uint64_t capacityVertices);

// billboards.cu: cl::BillboardSystem::renderMain: sort by (depth, index), cut, crop, expand
struct BillboardScratch {
	uint32_t *keys[2], *vals[2];     // [n] ping-pong pairs
	uint32_t *tileHist;              // [256 * ceil(n / 1024)]
	uint32_t *blockCounts;           // [ceil(min(n, budget) / 256)]
};
uint32_t launch_billboards(cudaStream_t st, const BillboardDev *bb, uint32_t n, uint32_t maxPerFrame, const BillboardScratch &scratch,
	float *verts, unsigned long long *atlasOut, uint32_t *totalOut /* [2]: quads, draws */, BillboardDrawDev *drawsOut /* room for min(n, maxPerFrame) */);

// sort.cu: segmented stable LSD radix sort of the (key, slot) pairs of the listed effects
// followed by the 4x expansion in sorted order.
struct SortScratch {
	uint32_t *keysAlt, *valsAlt;     // ping-pong buffers, same indexing as PoolPtrs::keys/vals
	uint32_t *hist;                  // [numSegments][4][256] digit histograms -> digit bases, then [numSegments] pass masks, then [4] pass tickets
	uint32_t *lookback;              // [4][numSortTiles][256] chained-scan status words
	bool deferExpansion;             // sprContextSetDeferredExpansion: sort, but leave the 4x expansion to the assembly
	uint32_t numSms, smallMaxSlots;  // k_sort_small: SMs of the device, slot bound of the largest one-CTA effect of the frame (for the key-range split)
};
// Optional timing callbacks around each sort kernel class (0 hist, 1 scan, 2 pass, 3 expand_sorted, 4 sort_small).
struct SortTimingHooks {
	void *ctx;
	void *(*begin)(void *ctx, int kernel);
	void (*end)(void *ctx, int kernel, void *token);
};
// Zeroes the histograms, pass masks and status words of one frame's sort. Issued at the START of the frame (before
// k_plan) so that no memset sits between the frame's kernels: the whole chain then launches programmatically (launch.h).
void clear_sort_scratch(cudaStream_t st, const SortScratch &scratch, uint32_t numSegments, uint32_t numTiles);
// segEffects / tiles list the tile-path ("big") effects first: [0, numBigSegments) / [0, numBigTiles) go through
// hist + scan + 4 passes, the effects behind them (at most kSmallSortMax slots each) through k_sort_small, one CTA
// per effect; k_expand_sorted then runs over all tiles. Returns the number of launches.
uint32_t launch_sort_and_expand(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, const SortScratch &scratch,
	const uint32_t *segEffects, uint32_t numSegments, uint32_t numBigSegments, const SortTile *tiles, uint32_t numTiles, uint32_t numBigTiles,
	const SortTimingHooks *hooks);

} // namespace spr
