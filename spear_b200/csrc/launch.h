// launch.h -- programmatic dependent launch (PDL) for the kernels of one frame.
//
// A frame is a chain of 12-20 dependent kernels, several of them a few microseconds long (k_plan, k_emit_warp,
// k_dead_append, k_finalize, k_sort_scan, a skipped k_sort_pass). Every kernel of the chain starts with pdl_enter():
// `griddepcontrol.launch_dependents` lets the NEXT kernel of the stream be launched and its CTAs become resident as
// soon as every CTA of this one has started, and `griddepcontrol.wait` then blocks until the PREVIOUS kernel has
// completed and its writes are visible -- so the data dependencies are exactly those of plain stream order, while
// launch latency, CTA scheduling and parameter loads overlap the predecessor's tail.
// Launched without the attribute (SPR_PDL=0, or after a memset / copy) both instructions are no-ops.
#pragma once

#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

namespace spr {

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_enter()
{
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	asm volatile("griddepcontrol.wait;" ::: "memory");
}
// The two halves of pdl_enter() for kernels that have work which depends on NOTHING the frame's earlier kernels write
// (shared-memory clears, reads of the work lists the host uploaded before the chain started): that work goes between
// pdl_launch_dependents() and pdl_wait() and overlaps the predecessor's tail. Everything a kernel of this frame may
// have written -- and with it every early exit that depends on such data -- stays behind pdl_wait(): a CTA ahead of the
// wait can be running while the predecessor of its predecessor is still at work (every kernel of the chain releases its
// successor when it STARTS). Only what reached the device with the frame's host-to-device copy (tile table, segment
// list, entry list) is read there: a kernel cannot be launched programmatically across a copy, so everything behind the
// copy starts after it is complete. Tables written by the upload kernels (params, granule owners) are read behind the wait.
#ifdef SPR_NO_HOIST // A/B build: nothing runs ahead of the wait
__device__ __forceinline__ void pdl_launch_dependents() { pdl_enter(); }
__device__ __forceinline__ void pdl_wait() { }
#else
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
#endif

inline bool pdl_enabled()
{
	static const bool on = [] { const char *e = getenv("SPR_PDL"); return !(e && e[0] == '0'); }();
	return on;
}

// kernel<<<grid, block, smem, st>>>(args...) with the programmatic-stream-serialization attribute
template <typename... KArgs, typename... Args>
inline void launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
	cudaLaunchAttribute attr;
	memset(&attr, 0, sizeof(attr));
	attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr.val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = &attr;
	cfg.numAttrs = pdl_enabled() ? 1u : 0u;
	cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

} // namespace spr
