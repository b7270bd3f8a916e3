// particle_system.h -- C++ host layer of the B200 particle path.
//
// spr::ParticleSystem mirrors the reference's operator interface
// (cl::ParticleSystem, src/client/ParticleSystem.h:9-22; implementation
// ParticleSystemImp, src/client/ParticleSystem.cpp:148-905): same entry points,
// same argument meaning, same id recycling, same (absent) error behaviour.
// spr::Context owns everything that lives on one GPU: the slot pool, the effect /
// type / system tables and the stream; any number of systems share one context so
// that a frame over thousands of independent systems is ONE set of kernel launches.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/spear_particles.h"
#include "device_types.h"
#include "host_math.h"
#include "host_pool.h"
#include "kernels.h"

namespace spr {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; }; // -> SPR_ERR_UNSUPPORTED

#define SPR_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) throw spr::CudaError(std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)

template <typename T>
struct DeviceArray {
	T *ptr = nullptr;
	size_t capacity = 0;
	~DeviceArray() { if (ptr) cudaFree(ptr); }
	// grow to at least n elements keeping the first `keep` elements; new tail zero-filled
	void reserve(size_t n, size_t keep, cudaStream_t st, int fillByte = 0) {
		if (n <= capacity) return;
		size_t cap = capacity ? capacity : 64;
		while (cap < n) cap *= 2;
		T *p = nullptr;
		SPR_CUDA(cudaMalloc(&p, cap * sizeof(T)));
		if (keep) SPR_CUDA(cudaMemcpyAsync(p, ptr, keep * sizeof(T), cudaMemcpyDeviceToDevice, st));
		SPR_CUDA(cudaMemsetAsync(p + keep, fillByte, (cap - keep) * sizeof(T), st));
		if (ptr) { SPR_CUDA(cudaStreamSynchronize(st)); cudaFree(ptr); }
		ptr = p; capacity = cap;
	}
};

template <typename T>
struct PinnedArray {
	T *ptr = nullptr;
	size_t capacity = 0;
	~PinnedArray() { if (ptr) cudaFreeHost(ptr); }
	void reserve(size_t n) {
		if (n <= capacity) return;
		size_t cap = capacity ? capacity : 64;
		while (cap < n) cap *= 2;
		if (ptr) cudaFreeHost(ptr);
		SPR_CUDA(cudaMallocHost(&ptr, cap * sizeof(T)));
		capacity = cap;
	}
};

// Host mirror of one effect (ParticleSystemImp::Effect, ParticleSystem.cpp:186-218), split in two parallel arrays.
// HostEffect holds what the per-frame pass of Context::update reads and writes for EVERY active effect (48 bytes: a
// 100k-effect frame streams 5 MB through the host instead of the 21 MB of the whole mirror); HostEffectCold holds the
// capacity bookkeeping (only effects that can still grow) and the transforms (render, updateTransform, localSpace bounds).
struct HostEffect {
	uint32_t typeId = 0;          // system-local
	uint32_t global = 0;          // context-wide effect index
	float timeDelta = 0.0f;
	float timeStep = 0.0f;
	double lastUpdateTime = 0.0;
	uint64_t simSteps = 0;
	uint64_t uploadFrameIndex = 16;
	bool firstEmit = true;        // host view: no sub-step simulated yet
	bool inUse = false;
	bool hostStopEmit = false;
	bool instantDelete = false;
	bool capacityFinal = false;   // the slot count provably never exceeds the bound fixed at addEffect: no per-frame capacity bookkeeping
};

struct HostEffectCold {
	uint32_t slotUpperBound = 0;  // conservative bound of particles.size*4
	// The device mirrors the exact slot count of every simulated effect to pinned host memory at the end of each
	// frame (Context::slotMirror). The bound is re-based on the newest mirrored count plus the growth allowance of
	// the frames submitted after it -- without that it only ever grows and a small long-lived effect would need a
	// synchronous read-back every few frames.
	uint32_t createdStamp = 0;    // frame stamp when the effect was added: older mirror entries belong to a previous owner of the index
	uint32_t evictedStamp = 0;    // newest frame whose allowance is no longer in `recent`
	uint32_t recentCount = 0;
	struct { uint32_t stamp, add; } recent[4] = { { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } }; // allowances of the last simulated frames
	SprVec3 origin = { 0, 0, 0 };
	SprMat34 particlesToWorld;
	SprMat34 boundsToWorld;       // particlesToWorld at the last simulated step (gpuBounds, :675-677)
};

struct HostType {
	const SprParticleSystemComponent *comp = nullptr; // pointer identity is the key (:30-41)
	uint32_t refCount = 0;
	uint32_t global = 0;          // context-wide type index
	double renderOrder = 0.0;
	uint32_t tiebreaker = 0;
	float timeStep = 0.0f;
	bool hasGravityPoints = false;
	SprVertexTypeUbo typeUbo;
	uint8_t lut[SPR_SPLINE_LUT_BYTES];
};

struct SortedEffect { // :225-239 (+ the particle count read back from the device, not part of the order)
	double renderOrder; uint32_t tiebreaker; float depth; uint32_t effectId; uint32_t numParticles;
	bool operator<(const SortedEffect &rhs) const {
		if (renderOrder != rhs.renderOrder) return renderOrder < rhs.renderOrder;
		if (tiebreaker != rhs.tiebreaker) return tiebreaker < rhs.tiebreaker;
		if (depth != rhs.depth) return depth < rhs.depth;
		if (effectId != rhs.effectId) return effectId < rhs.effectId;
		return false;
	}
};

struct Context;

struct ParticleSystem {
	Context *ctx;
	uint32_t systemIdx;           // context-wide
	HostRng initRng;              // :250
	std::vector<HostEffect> effects;          // :241 (hot half)
	std::vector<HostEffectCold> effectsCold;  // :241 (cold half, same index)
	std::vector<uint32_t> freeEffectIds;      // :242
	// per effect id: the mark of the last update call that listed it (kMarkFree while the id is not in use). One 4-byte
	// look-up per listed id validates the id and finds an id listed twice, before the call changes anything.
	static constexpr uint32_t kMarkFree = 0xffffffffu;
	std::vector<uint32_t> activeMark;
	uint32_t systemMark = 0;                  // the mark of the last update call that listed this system
	std::vector<HostType> types;              // :244
	std::vector<uint32_t> freeTypeIds;        // :245
	std::unordered_map<const void*, uint32_t> componentToType; // :246
	uint64_t bufferFrameIndex;                // :263
	std::vector<SprDrawRecord> drawRecords;
	std::vector<SortedEffect> sortedScratch;  // renderMain's sort array, kept between frames
	uint64_t renderBatchStamp = 0;            // Context::renderBatch: the last threaded batch that named this system

	ParticleSystem(Context *ctx, const SprSystemsDesc &desc); // :265-267
	~ParticleSystem();

	uint32_t reserveEffectType(const SprParticleSystemComponent *c);   // :696-720
	void releaseEffectType(const SprParticleSystemComponent *c);       // :722-728
	uint32_t addEffect(uint32_t entityId, uint8_t componentIndex, const SprParticleSystemComponent *c, const SprTransform &transform); // :730-766
	void updateTransform(uint32_t effectId, const SprTransformUpdate &update); // :768-784
	bool prepareForRemove(uint32_t effectId, const SprFrameArgs &args);        // :786-794
	void remove(uint32_t effectId);                                            // :796-806
	void updateParticles(const uint32_t *activeIds, uint32_t numActive, const SprFrameArgs &args); // :808-822
	void renderMain(const uint32_t *visibleIds, uint32_t numVisible, const SprRenderArgs &args, SprRenderOutput &out); // :824-903

	// effects whose sphere area intersects the frustum, ascending ids (AreaSystem::queryFrustum over the particle spheres); returns the count
	uint32_t queryEffects(const SprFrustum &frustum, uint32_t *outIds, uint32_t capacity);
	// Particle.glsl `vs` over draw records (vertex_stage.cu); returns the number of vertices written
	uint64_t shadeVertices(const SprDrawRecord *records, uint32_t numRecords, const SprFogImage *fog, SprShadedVertex *deviceOut, uint64_t capacityVertices);
	void renderMainWith(const uint32_t *visibleIds, uint32_t numVisible, const SprRenderArgs &args, const RenderGather *got, SprRenderOutput &out);
	void releaseEffectTypeImp(uint32_t typeId);                        // :682-692
};

SprBounds3 boundsFromGather(const RenderGather &g, const HostType &type, const HostEffectCold &C); // gpuBounds, :672-677

struct UpdateRequest {
	ParticleSystem *system;
	const uint32_t *activeIds;
	uint32_t numActive;
	const SprFrameArgs *args;
};

struct Context {
	int device = 0;
	uint32_t flags = 0;
	uint32_t maxParticlesPerFrame = 4096;
	uint32_t maxParticleCacheFrames = 16;
	cudaStream_t stream = nullptr;
	bool ownStream = false;
	uint32_t numSms = 148;
	bool deferExpansion = false;               // sprContextSetDeferredExpansion (sort mode): updates sort but do not write the local vertex stream
	uint32_t assemblyKernel = 0;               // sprContextSetAssemblyKernel: 0 = by share, 1 = load/store kernel, 2 = bulk-asynchronous pipeline
	uint32_t assemblyCtasPerSm = 8;            // sprContextSetAssemblyShare: resident CTAs per SM of sprExpandFromPeers
	uint64_t launchCount = 0;
	uint64_t totalSimSteps = 0;            // simulateParticlesImp calls issued so far (all effects)
	std::vector<uint8_t> effectLive;       // per effect global: currently in use

	// ---- slot pool
	uint64_t poolCapacity = 0;     // slots
	uint64_t poolUsed = 0;         // bump pointer
	uint64_t poolOwnedEnd = 0;     // end of the highest slot range any effect owns granules in: the per-frame kernels stop there
	float4 *dSlotState = nullptr;
	uint32_t *dFreeStack = nullptr;
	float4 *dVertices = nullptr;
	uint32_t *dKeys[2] = { nullptr, nullptr };
	uint32_t *dVals[2] = { nullptr, nullptr };
	uint32_t *dAliveMask = nullptr, *dDeadMask = nullptr, *dWordAliveBase = nullptr;
	uint32_t *dGranuleOwner = nullptr;
	unsigned long long *dTileStatus[2] = { nullptr, nullptr };
	unsigned long long *dTickets = nullptr;    // [2] monotonic tile tickets (PoolPtrs::tickets) ...
	uint64_t ticketBase[2] = { 0, 0 };         // ... and how many of them the launches issued so far have taken
	std::vector<std::vector<uint32_t>> freeRanges; // per log2(capacity): free slotBase values

	// ---- tables
	DeviceArray<EffectParams> dParams;
	DeviceArray<EffectState> dState;
	DeviceArray<RoundConsts> dRoundConsts;
	DeviceArray<uint32_t> dEffectAnyDead;
	DeviceArray<TypeDev> dTypes;
	DeviceArray<SystemDev> dSystems;
	DeviceArray<PlanRec> dPlans;
	std::vector<EffectParams> hParams;    // host copy, indexed by global effect index
	std::vector<uint32_t> ownedSlots;     // per effect: slots [0, ownedSlots) have their granuleOwner entries set
	std::vector<uint32_t> slotCapacityOf; // per effect: copy of hParams[g].slotCapacity, dense (the per-frame pass reads only this)
	std::vector<uint32_t> freeEffectGlobals, freeTypeGlobals, freeSystemGlobals;
	uint32_t numEffectGlobals = 0, numTypeGlobals = 0, numSystemGlobals = 0;

	// ---- pending uploads (applied at the next submit / read-back)
	std::vector<uint32_t> dirtyParams;        // global effect indices
	std::vector<uint8_t> dirtyParamsFlag;
	std::vector<uint32_t> initStateIdx; std::vector<EffectState> initStateVal;
	std::vector<uint32_t> typeIdx; std::vector<TypeDev> typeVal;
	std::vector<uint32_t> systemIdx; std::vector<SystemDev> systemVal;
	std::vector<FillJob> ownerJobs;
	// One launch scatters a whole pending list in parallel, so an index may appear in it only once: a
	// recycled index (type / effect / system removed and re-added before the next submit) replaces its
	// pending value instead of being queued a second time. pos[i] = 1 + position of i in the list.
	std::vector<uint32_t> initStatePos, typePos, systemPos;
	template <typename T>
	static void queueUpload(std::vector<uint32_t> &idx, std::vector<T> &val, std::vector<uint32_t> &pos, uint32_t i, const T &v)
	{
		if (pos.size() <= i) pos.resize((size_t)i + 1, 0);
		if (pos[i]) { val[pos[i] - 1] = v; return; }
		idx.push_back(i); val.push_back(v);
		pos[i] = (uint32_t)idx.size();
	}
	static void clearUploads(std::vector<uint32_t> &idx, std::vector<uint32_t> &pos) { for (uint32_t i : idx) pos[i] = 0; idx.clear(); }

	// ---- per-frame work lists of Context::update (kept for their capacity)
	std::vector<ActiveEntry> scratchEntries;
	std::vector<SystemWork> scratchWork;
	std::vector<BurstJob> scratchBurstJobs;
	std::vector<uint32_t> scratchSegEffects, scratchSmallEffects;
	std::vector<SortTile> scratchSortTiles;
	std::vector<EffectParams> scratchParamVals;
	int updateDepth = 0;
	uint32_t updateMark = 0;               // numbers the update calls (ParticleSystem::activeMark / systemMark)
	std::vector<ParticleSystem*> systemsAlive; // every system of the context (marks are wiped when updateMark wraps)
	void validateRequests(const UpdateRequest *reqs, uint32_t numReqs);
	HostPool hostPool;                     // persistent workers for the large per-system / per-effect host passes
	uint32_t hostPassThreads(size_t totalActive, uint32_t numReqs) const;
	static std::vector<uint32_t> splitRequests(const UpdateRequest *reqs, uint32_t numReqs, uint32_t nt);
	uint64_t renderBatchStamp = 0;     // numbers the threaded render batches (a system named twice keeps the batch on one thread)
	// ---- per-frame staging
	PinnedArray<uint8_t> stagingH[2];
	DeviceArray<uint8_t> stagingD;
	cudaEvent_t stagingEvent[2] = { nullptr, nullptr };
	int stagingParity = 0;
	PinnedArray<RenderGather> gatherH;
	PinnedArray<uint32_t> gatherIdxH;
	DeviceArray<RenderGather> gatherD;
	DeviceArray<uint32_t> gatherIdxD;
	int streamFusedCtasPerSm[2] = { 1, 1 };
	VertexTensorMap vertexMap;                 // tensor map of dVertices for k_stream_fused's bulk stores (re-encoded when the pool grows)
	bool vertexMapValid = false;
	bool streamTma = true;                     // SPR_STREAM_TMA=0: expansion stores through the LSU (A/B)
	bool streamFused = true;               // stream mode: one-pass pipelined kernel (default) or integrate + dead_append + expand
	uint32_t numTinyEffects = 0;          // live effects whose whole range is one 128-slot granule
	uint32_t numLiveEffects = 0;          // live effects of every size
	uint32_t typesWithGravityPoints = 0; // live effect types with gravityPoints (selects the kernel variant)
	uint32_t frameStamp = 0;
	uint32_t epoch = 0;

	// ---- optional per-kernel timing (bench.py roofline): CUDA events around every launch class
	struct TimedLaunch { int kernel; cudaEvent_t begin, end; };
	bool timingEnabled = false;
	std::vector<TimedLaunch> timedLaunches;
	std::vector<cudaEvent_t> eventPool;
	double kernelMs[17] = { 0 };
	uint64_t kernelLaunches[17] = { 0 };
	cudaEvent_t timingBegin(int kernel);
	void timingEnd(int kernel, cudaEvent_t begin);
	void resolveTimings();

	// ---- sort scratch
	DeviceArray<uint32_t> dSortHist, dSortLookback;
	unsigned long long *slotMirror = nullptr; // pinned + mapped, one word per effect: frame stamp << 32 | slot count (written by k_finalize)
	size_t slotMirrorCapacity = 0;
	void ensureSlotMirror(size_t numEffects);
	DeviceArray<uint32_t> queryD;         // [count, ids...] of the last sprQueryEffects
	DeviceArray<uint8_t> shadeBlobD;      // job table + LUT table of the last sprShadeVertices
	DeviceArray<uint64_t> dRunOffsets;
	DeviceArray<uint32_t> packListD;      // effect list of the last sprRunBufferPack (device)
	std::vector<uint32_t> packList;       // ... and its host copy

	explicit Context(const SprContextDesc &desc);
	~Context();

	PoolPtrs poolPtrs() const;
	TablePtrs tablePtrs();

	uint32_t allocEffectGlobal();
	uint32_t allocRange(uint32_t capacity);
	void freeRange(uint32_t slotBase, uint32_t capacity);
	void ensurePool(uint64_t slots);
	void markParamsDirty(uint32_t g);
	void ensureOwned(uint32_t g, uint32_t slotUpperBound);
	void flushUploads();
	void update(const UpdateRequest *reqs, uint32_t numReqs);
	void gatherEffects(const uint32_t *globals, uint32_t n, RenderGather *out); // synchronous read-back
	uint32_t *gatherIndexBuffer(uint32_t n);           // pinned: the caller writes n context-wide effect indices here ...
	const RenderGather *gatherSubmit(uint32_t n);      // ... and gets their records back in pinned memory (valid until the next gather)
	void renderBatch(ParticleSystem *const *systems, uint32_t numSystems, const uint32_t *const *visibleIds,
		const uint32_t *numVisible, const SprRenderArgs *args, size_t argsStride, SprRenderOutput *outs);
	void growEffect(uint32_t g, uint32_t minCapacity, bool copyContents);
	void sync() { SPR_CUDA(cudaStreamSynchronize(stream)); }
};

// cl::BillboardSystem (src/client/BillboardSystem.{h,cpp}) over billboards.cu: SURVEY.md 8(f) row N4.
struct BillboardSystem {
	Context *ctx;
	uint32_t maxPerFrame;                      // MaxBillboardsPerFrame, BillboardSystem.cpp:10
	BillboardDev *pending = nullptr;           // billboards + billboardOrder of the frame (:61-62) that came through add(), in pinned host memory
	size_t pendingCount = 0, pendingCapacity = 0;
	// The frame in call order: runs of converted billboards (in `pending`) and packed arrays of the caller (host or device)
	struct Segment { int kind; const BillboardDev *src; size_t offset; size_t count; }; // kind 0: pending + offset, 1: caller's host array src, 2: caller's device array src
	std::vector<Segment> segments;
	size_t frameCount = 0;
	void addPacked(const BillboardDev *packed, uint32_t count, bool onDevice);
	static BillboardDev *pack(const SprBillboard *billboards, uint32_t i0, uint32_t i1, BillboardDev *dst);
	~BillboardSystem() { if (pending) cudaFreeHost(pending); }
	PinnedArray<SprBillboardDraw> drawsPinned; // the frame's draw list, built on the device (k_bb_run_*) and read back
	uint32_t lastQuads = 0;                    // quads of the last renderMain (sprBillboardReadVertices)
	DeviceArray<BillboardDev> dBillboards;
	DeviceArray<uint32_t> dKeys[2], dVals[2], dTileHist, dBlockCounts, dTotal;
	DeviceArray<float> dVerts;                 // billboardVertices (:65): 24 words per quad
	DeviceArray<unsigned long long> dAtlas;
	DeviceArray<BillboardDrawDev> dDraws;

	BillboardSystem(Context *c, uint32_t maxBillboardsPerFrame);
	void add(const SprBillboard *billboards, uint32_t count);                 // :80-117
	void renderMain(const SprRenderArgs &args, SprBillboardOutput &out);      // :119-228
};

// descriptor_json.cpp: effect descriptors from the reference's JSON prefab format
struct DescriptorSet;
DescriptorSet *descriptorsFromJson(const char *json, size_t length, std::string &error);
uint32_t descriptorCount(const DescriptorSet *s);
const SprParticleSystemComponent *descriptorGet(const DescriptorSet *s, uint32_t i);
const char *descriptorTexture(const DescriptorSet *s, uint32_t i);
void descriptorsFree(DescriptorSet *s);

} // namespace spr
