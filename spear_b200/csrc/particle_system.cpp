// particle_system.cpp -- host layer (see particle_system.h).
//
// The host keeps only what the reference's callers can observe without running the
// simulation (ids, type tables, timeDelta / sub-step counts, transforms); everything
// a sub-step reads or writes (timers, RNG, slots, free lists, streams, bounds) lives
// on the device. There is no CPU simulation path in this file.

#include "particle_system.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <exception>
#include <mutex>
#include <thread>
#include <chrono>
#include <map>
#include <stdio.h>
#include <unordered_map>
#include <stdlib.h>
#include <string.h>

namespace spr {

static uint32_t pow2ceil(uint64_t v)
{
	uint64_t p = kGranuleSlots;
	while (p < v) p <<= 1;
	if (p > 0x80000000ull) throw std::runtime_error("effect larger than 2^31 slots");
	return (uint32_t)p;
}

static uint32_t log2u(uint32_t v) { uint32_t l = 0; while ((1u << l) < v) l++; return l; }

// ---------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------

Context::Context(const SprContextDesc &desc)
{
	int count = 0;
	cudaError_t err = cudaGetDeviceCount(&count);
	if (err != cudaSuccess || count == 0) throw CudaError(std::string("no CUDA device: ") + cudaGetErrorString(err));
	device = desc.device;
	SPR_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	SPR_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) throw CudaError("spear_particles kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));
	numSms = (uint32_t)prop.multiProcessorCount;
	flags = desc.flags;
	maxParticlesPerFrame = desc.maxParticlesPerFrame ? desc.maxParticlesPerFrame : 4096;
	maxParticleCacheFrames = desc.maxParticleCacheFrames ? desc.maxParticleCacheFrames : 16;
	if (desc.stream) { stream = (cudaStream_t)desc.stream; ownStream = false; }
	else { SPR_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)); ownStream = true; }
	SPR_CUDA(cudaEventCreateWithFlags(&stagingEvent[0], cudaEventDisableTiming));
	SPR_CUDA(cudaEventCreateWithFlags(&stagingEvent[1], cudaEventDisableTiming));
	streamFused = getenv("SPR_STREAM_TWO_PASS") == nullptr; // SPR_STREAM_TWO_PASS=1 selects the two-pass stream mode (A/B measurement)
	configure_stream_fused(streamFusedCtasPerSm);
	{ const char *e = getenv("SPR_STREAM_TMA"); streamTma = !(e && e[0] == '0'); }
	configure_sort_kernels();
	freeRanges.resize(33);
	SPR_CUDA(cudaMalloc(&dTickets, 2 * sizeof(unsigned long long)));
	SPR_CUDA(cudaMemsetAsync(dTickets, 0, 2 * sizeof(unsigned long long), stream));
	ensurePool(desc.initialSlotCapacity ? desc.initialSlotCapacity : (1u << 20));
}

#ifdef SPR_SORT_PROFILE
void sort_profile_dump(); // sort.cu, instrumented experiment builds only (SPR_EXTRA_NVCC_FLAGS=-DSPR_SORT_PROFILE)
#endif

Context::~Context()
{
	cudaSetDevice(device);
	cudaStreamSynchronize(stream);
#ifdef SPR_SORT_PROFILE
	sort_profile_dump();
#endif
	cudaFree(dSlotState); cudaFree(dFreeStack); cudaFree(dVertices);
	cudaFree(dKeys[0]); cudaFree(dKeys[1]); cudaFree(dVals[0]); cudaFree(dVals[1]);
	cudaFree(dAliveMask); cudaFree(dDeadMask); cudaFree(dWordAliveBase);
	cudaFree(dGranuleOwner); cudaFree(dTileStatus[0]); cudaFree(dTileStatus[1]); cudaFree(dTickets);
	if (slotMirror) cudaFreeHost(slotMirror);
	cudaEventDestroy(stagingEvent[0]); cudaEventDestroy(stagingEvent[1]);
	if (ownStream) cudaStreamDestroy(stream);
}

PoolPtrs Context::poolPtrs() const
{
	PoolPtrs p;
	p.state = dSlotState; p.freeStack = dFreeStack; p.vertices = dVertices;
	p.keys = dKeys[0]; p.vals = dVals[0];
	p.aliveMask = dAliveMask; p.deadMask = dDeadMask; p.wordAliveBase = dWordAliveBase;
	p.granuleOwner = dGranuleOwner;
	p.tileStatusAlive = dTileStatus[0]; p.tileStatusDead = dTileStatus[1];
	p.tickets = dTickets;
	return p;
}

TablePtrs Context::tablePtrs()
{
	TablePtrs t;
	t.params = dParams.ptr; t.state = dState.ptr; t.types = dTypes.ptr; t.systems = dSystems.ptr;
	t.plans = dPlans.ptr;
	t.roundConsts = dRoundConsts.ptr;
	t.effectAnyDead = dEffectAnyDead.ptr;
	t.slotMirror = slotMirror; // unified addressing: the pinned allocation is reachable from the device under the same pointer
	return t;
}

template <typename T>
static void regrow(T *&ptr, uint64_t oldCount, uint64_t newCount, uint64_t keepCount, int fillByte, cudaStream_t st)
{
	T *p = nullptr;
	SPR_CUDA(cudaMalloc(&p, newCount * sizeof(T)));
	if (keepCount) SPR_CUDA(cudaMemcpyAsync(p, ptr, keepCount * sizeof(T), cudaMemcpyDeviceToDevice, st));
	if (fillByte >= 0) SPR_CUDA(cudaMemsetAsync(p + keepCount, fillByte, (newCount - keepCount) * sizeof(T), st));
	(void)oldCount;
	if (ptr) { SPR_CUDA(cudaStreamSynchronize(st)); SPR_CUDA(cudaFree(ptr)); }
	ptr = p;
}

void Context::ensurePool(uint64_t slots)
{
	if (slots <= poolCapacity) return;
	uint64_t cap = poolCapacity ? poolCapacity * 2 : kTileSlots;
	while (cap < slots) cap *= 2;
	if (cap > 0xffffffffull) throw std::runtime_error("slot pool larger than 2^32 slots");
	uint64_t used = poolCapacity; // copy everything that was addressable
	regrow(dSlotState, poolCapacity * 2, cap * 2, used * 2, -1, stream);
	regrow(dFreeStack, poolCapacity, cap, used, -1, stream);
	regrow(dVertices, poolCapacity * 8, cap * 8, used * 8, -1, stream);
	regrow(dGranuleOwner, poolCapacity / kGranuleSlots, cap / kGranuleSlots, used / kGranuleSlots, 0xff, stream);
	regrow(dTileStatus[0], poolCapacity / kTileSlots, cap / kTileSlots, 0, 0, stream);
	regrow(dTileStatus[1], poolCapacity / kTileSlots, cap / kTileSlots, 0, 0, stream);
	if (flags & SPR_CTX_SORT_PARTICLES) {
		for (int i = 0; i < 2; i++) {
			regrow(dKeys[i], poolCapacity, cap, 0, -1, stream);
			regrow(dVals[i], poolCapacity, cap, i == 0 ? used : 0, -1, stream); // vals[0] holds the sorted slot order of every effect
		}
	}
	{
		// the ballots (and, in stream mode, the word ranks) are rewritten by every simulated step, but an
		// effect that is not simulated in the frame after a pool growth still needs its old ones
		regrow(dAliveMask, poolCapacity / 32, cap / 32, poolCapacity / 32, 0, stream);
		regrow(dDeadMask, poolCapacity / 32, cap / 32, poolCapacity / 32, 0, stream);
		if (!(flags & SPR_CTX_SORT_PARTICLES)) regrow(dWordAliveBase, poolCapacity / 32, cap / 32, poolCapacity / 32, 0, stream);
	}
	poolCapacity = cap;
	vertexMapValid = streamTma && !(flags & SPR_CTX_SORT_PARTICLES) && encode_vertex_map(&vertexMap, dVertices, poolCapacity);
}

uint32_t Context::allocRange(uint32_t capacity)
{
	uint32_t cls = log2u(capacity);
	if (!freeRanges[cls].empty()) {
		uint32_t base = freeRanges[cls].back();
		freeRanges[cls].pop_back();
		return base;
	}
	uint64_t align = capacity < kTileSlots ? capacity : kTileSlots;
	uint64_t base = (poolUsed + align - 1) / align * align;
	poolUsed = base + capacity;
	ensurePool(poolUsed);
	return (uint32_t)base;
}

void Context::freeRange(uint32_t slotBase, uint32_t capacity)
{
	freeRanges[log2u(capacity)].push_back(slotBase);
	ownerJobs.push_back({ slotBase / kGranuleSlots, capacity / kGranuleSlots, kNoOwner, 0 });
}

uint32_t Context::allocEffectGlobal()
{
	uint32_t g;
	if (!freeEffectGlobals.empty()) { g = freeEffectGlobals.back(); freeEffectGlobals.pop_back(); }
	else {
		g = numEffectGlobals++;
		hParams.resize(numEffectGlobals);
		ownedSlots.resize(numEffectGlobals, 0);
		slotCapacityOf.resize(numEffectGlobals, 0);
		effectLive.resize(numEffectGlobals, 0);
		dirtyParamsFlag.resize(numEffectGlobals, 0);
		size_t keep = numEffectGlobals - 1;
		dParams.reserve(numEffectGlobals, keep, stream);
		dState.reserve(numEffectGlobals, keep, stream);
		dRoundConsts.reserve(numEffectGlobals, keep, stream);
		dEffectAnyDead.reserve(numEffectGlobals, keep, stream);
		ensureSlotMirror(numEffectGlobals);
	}
	effectLive[g] = 1;
	numLiveEffects++;
	return g;
}

void Context::ensureSlotMirror(size_t numEffects)
{
	if (numEffects <= slotMirrorCapacity) return;
	size_t cap = slotMirrorCapacity ? slotMirrorCapacity * 2 : 1024;
	while (cap < numEffects) cap *= 2;
	unsigned long long *p = nullptr;
	SPR_CUDA(cudaHostAlloc(&p, cap * sizeof(unsigned long long), cudaHostAllocMapped));
	memset(p, 0, cap * sizeof(unsigned long long));
	if (slotMirror) {
		SPR_CUDA(cudaStreamSynchronize(stream)); // frames in flight still write the old array
		memcpy(p, slotMirror, slotMirrorCapacity * sizeof(unsigned long long));
		cudaFreeHost(slotMirror);
	}
	slotMirror = p; slotMirrorCapacity = cap;
}

void Context::markParamsDirty(uint32_t g)
{
	if (!dirtyParamsFlag[g]) { dirtyParamsFlag[g] = 1; dirtyParams.push_back(g); }
}

void Context::growEffect(uint32_t g, uint32_t minCapacity, bool copyContents)
{
	EffectParams &P = hParams[g];
	uint32_t newCap = pow2ceil(minCapacity);
	if (newCap <= P.slotCapacity) return;
	uint32_t oldBase = P.slotBase, oldCap = P.slotCapacity;
	uint32_t newBase = allocRange(newCap);
	if (copyContents && oldCap) {
		SPR_CUDA(cudaMemcpyAsync(dSlotState + 2ull * newBase, dSlotState + 2ull * oldBase, (size_t)oldCap * 32, cudaMemcpyDeviceToDevice, stream));
		SPR_CUDA(cudaMemcpyAsync(dFreeStack + newBase, dFreeStack + oldBase, (size_t)oldCap * 4, cudaMemcpyDeviceToDevice, stream));
	}
	if (oldCap) freeRange(oldBase, oldCap);
	if (oldCap == kGranuleSlots) numTinyEffects--;
	if (newCap == kGranuleSlots) numTinyEffects++;
	uint32_t owned = ownedSlots[g];
	ownedSlots[g] = 0;
	P.slotBase = newBase;
	P.slotCapacity = newCap;
	slotCapacityOf[g] = newCap;
	ensureOwned(g, owned);
	markParamsDirty(g);
}

// Granule ownership is only set for the part of the range the effect can have reached
// (its conservative slot bound, rounded up to a tile): the streaming kernels issue their record
// loads as soon as they know a granule is owned, so owned-but-unused granules would be wasted reads.
void Context::ensureOwned(uint32_t g, uint32_t slotUpperBound)
{
	uint64_t target = ((uint64_t)slotUpperBound + kTileSlots - 1) / kTileSlots * kTileSlots;
	if (target > slotCapacityOf[g]) target = slotCapacityOf[g];
	if (target <= ownedSlots[g]) return;          // the common case, decided from the two dense per-effect arrays
	const EffectParams &P = hParams[g];
	ownerJobs.push_back({ (P.slotBase + ownedSlots[g]) / kGranuleSlots, (uint32_t)(target - ownedSlots[g]) / kGranuleSlots, g, 0 });
	ownedSlots[g] = (uint32_t)target;
	if ((uint64_t)P.slotBase + target > poolOwnedEnd) poolOwnedEnd = (uint64_t)P.slotBase + target;
}

// Kernel classes for the optional timing hooks (names in c_api.cpp: sprKernelName).
enum { kKScatter = 0, kKPlan, kKEmit, kKEmitBurst, kKIntegrate, kKFinalize, kKSort, kKGather, kKPack, kKExpandRuns,
	kKSortBase = 10 /* + 0 hist, 1 scan, 2 pass, 3 expand_sorted */, kKDeadAppend = 14, kKExpandStream = 15, kKStreamFused = 6, kKSortSmall = 16 };

cudaEvent_t Context::timingBegin(int kernel)
{
	(void)kernel;
	if (!timingEnabled) return nullptr;
	cudaEvent_t e;
	if (!eventPool.empty()) { e = eventPool.back(); eventPool.pop_back(); }
	else SPR_CUDA(cudaEventCreate(&e));
	SPR_CUDA(cudaEventRecord(e, stream));
	return e;
}

void Context::timingEnd(int kernel, cudaEvent_t begin)
{
	if (!timingEnabled || !begin) return;
	cudaEvent_t e;
	if (!eventPool.empty()) { e = eventPool.back(); eventPool.pop_back(); }
	else SPR_CUDA(cudaEventCreate(&e));
	SPR_CUDA(cudaEventRecord(e, stream));
	timedLaunches.push_back({ kernel, begin, e });
}

void Context::resolveTimings()
{
	SPR_CUDA(cudaStreamSynchronize(stream));
	for (const TimedLaunch &t : timedLaunches) {
		float ms = 0.0f;
		SPR_CUDA(cudaEventElapsedTime(&ms, t.begin, t.end));
		kernelMs[t.kernel] += ms;
		kernelLaunches[t.kernel]++;
		eventPool.push_back(t.begin);
		eventPool.push_back(t.end);
	}
	timedLaunches.clear();
}

// SPR_DEBUG_SYNC=1: synchronise after every launch and name the kernel that faulted (debug aid only).
static const bool kDebugSync = getenv("SPR_DEBUG_SYNC") != nullptr;
#define TIMED(kernel, call) do { cudaEvent_t tb_ = timingBegin(kernel); call; launchCount++; timingEnd(kernel, tb_); \
	if (kDebugSync) { cudaError_t de_ = cudaStreamSynchronize(stream); if (de_ != cudaSuccess) throw CudaError(std::string(#call) + " -> " + cudaGetErrorString(de_)); } } while (0)

// One staging blob per submit: pending table uploads + the frame's work lists.
namespace {
struct BlobWriter {
	uint8_t *host; size_t size = 0;
	explicit BlobWriter(uint8_t *h) : host(h) { }
	template <typename T> size_t put(const T *src, size_t n) {
		size = (size + 15) & ~(size_t)15;
		size_t off = size;
		if (n) memcpy(host + off, src, n * sizeof(T));
		size += n * sizeof(T);
		return off;
	}
};
}

void Context::flushUploads()
{
	update(nullptr, 0); // an empty frame: only the pending table uploads are applied
}

// Argument check of an update call, before anything is changed: every id names a live effect, no id is listed twice
// in one system's list, no system is listed twice. The reference simulates a repeated id twice (its loop runs per list
// entry, ParticleSystem.cpp:812-821); the device plan works per effect, so that case is refused (SPR_ERR_UNSUPPORTED)
// instead of silently diverging.
void Context::validateRequests(const UpdateRequest *reqs, uint32_t numReqs)
{
	if (++updateMark >= ParticleSystem::kMarkFree - 1) { // the 32-bit mark wrapped: forget every old mark
		for (ParticleSystem *s : systemsAlive) { s->systemMark = 0; for (uint32_t &m : s->activeMark) if (m != ParticleSystem::kMarkFree) m = 0; }
		updateMark = 1;
	}
	const uint32_t mark = updateMark;
	size_t totalActive = 0;
	for (uint32_t r = 0; r < numReqs; r++) {
		ParticleSystem *sys = reqs[r].system;
		if (sys->systemMark == mark) throw Unsupported("a system is listed twice in one batched update (its effects share one RNG stream: merge the lists)");
		sys->systemMark = mark;
		if (reqs[r].numActive && !reqs[r].activeIds) throw std::invalid_argument("active id list is NULL");
		totalActive += reqs[r].numActive;
	}
	// the id marks belong to one system each: a large frame checks its lists on several host threads
	auto checkIds = [&](uint32_t r0, uint32_t r1) {
		for (uint32_t r = r0; r < r1; r++) {
			ParticleSystem *sys = reqs[r].system;
			uint32_t *marks = sys->activeMark.data();
			const size_t n = sys->activeMark.size();
			for (uint32_t i = 0; i < reqs[r].numActive; i++) {
				const uint32_t id = reqs[r].activeIds[i];
				if (id >= n || marks[id] == ParticleSystem::kMarkFree) throw std::invalid_argument("active list names an effect id that is not in use");
				if (marks[id] == mark) throw Unsupported("an effect id is listed twice in one active list (the reference would simulate it twice, ParticleSystem.cpp:812-821; not supported)");
				marks[id] = mark;
			}
		}
	};
	const uint32_t nt = hostPassThreads(totalActive, numReqs);
	if (nt <= 1) { checkIds(0, numReqs); return; }
	std::vector<uint32_t> cut = splitRequests(reqs, numReqs, nt);
	hostPool.run(nt, [&](uint32_t t) { checkIds(cut[t], cut[t + 1]); });
}

// Host threads for a per-effect pass over `totalActive` listed effects in `numReqs` systems (1 = the calling thread).
uint32_t Context::hostPassThreads(size_t totalActive, uint32_t numReqs) const
{
	if (totalActive < 16384 || numReqs < 16) return 1;
	const uint32_t most = HostPool::maxThreads() < 4 ? HostPool::maxThreads() : 4; // the pass streams ~60 bytes per effect: four cores are at memory speed
	return most < numReqs ? most : numReqs;
}

// Cuts the request list into nt contiguous ranges of about the same number of listed effects.
std::vector<uint32_t> Context::splitRequests(const UpdateRequest *reqs, uint32_t numReqs, uint32_t nt)
{
	size_t total = 0;
	for (uint32_t r = 0; r < numReqs; r++) total += reqs[r].numActive;
	std::vector<uint32_t> cut(nt + 1, numReqs);
	cut[0] = 0;
	size_t run = 0;
	uint32_t t = 1;
	for (uint32_t r = 0; r < numReqs && t < nt; r++) {
		run += reqs[r].numActive;
		while (t < nt && run >= total * t / nt) cut[t++] = r + 1;
	}
	return cut;
}

void Context::update(const UpdateRequest *reqs, uint32_t numReqs)
{
	SPR_CUDA(cudaSetDevice(device));
	static const bool hostTiming = getenv("SPR_HOST_TIMING") != nullptr; // debug aid: where a frame's host time goes
	auto tNow = [] { return std::chrono::steady_clock::now(); };
	auto tv = tNow();
	validateRequests(reqs, numReqs);
	const bool sortMode = (flags & SPR_CTX_SORT_PARTICLES) != 0;
	auto t0 = tNow();

	// ---- host pass: sub-step counts (ParticleSystem.cpp:810-820) and capacity bounds
	size_t totalActive = 0;
	for (uint32_t r = 0; r < numReqs; r++) totalActive += reqs[r].numActive;
	// the frame's work lists live in the context between frames: a 100k-effect frame would otherwise allocate, fault in
	// and release megabytes of vectors every call
	// (update re-enters itself with an empty request list when a read-back in the middle of the host pass has to flush
	// pending uploads, flushUploads(): the nested call gets lists of its own)
	struct DepthGuard { int &d; DepthGuard(int &x) : d(x) { d++; } ~DepthGuard() { d--; } } depthGuard(updateDepth);
	const bool nested = updateDepth > 1;
	std::vector<ActiveEntry> nestedEntries; std::vector<SystemWork> nestedWork; std::vector<BurstJob> nestedBurstJobs;
	std::vector<uint32_t> nestedSeg, nestedSmall; std::vector<SortTile> nestedTiles; std::vector<EffectParams> nestedParamVals;
	std::vector<ActiveEntry> &entries = nested ? nestedEntries : scratchEntries;
	std::vector<SystemWork> &work = nested ? nestedWork : scratchWork;
	std::vector<BurstJob> &burstJobs = nested ? nestedBurstJobs : scratchBurstJobs;
	entries.clear(); work.clear(); burstJobs.clear();
	entries.reserve(totalActive);
	work.reserve(numReqs);
	struct Resolve { HostEffect *E; HostEffectCold *C; uint32_t add; };
	std::vector<Resolve> resolve;
	std::vector<uint32_t> resolveGlobals;
	uint32_t numPlans = 0, maxSub = 0, burstBlocks = 0;
	const uint32_t thisStamp = frameStamp + 1; // the stamp this frame gets if it simulates anything
	entries.resize(totalActive);
	work.resize(numReqs);
	{
		uint32_t first = 0;
		for (uint32_t r = 0; r < numReqs; r++) {
			// every system is sorted towards the camera of its own request (a batch is equivalent to the per-system calls)
			const SprVec3 &cam = reqs[r].args->cameraPosition;
			work[r] = { reqs[r].system->systemIdx, first, reqs[r].numActive, 0, { cam.x, cam.y, cam.z }, 0 };
			first += reqs[r].numActive;
		}
	}

	// Phase 1, per effect and touching nothing but the effect itself: timeDelta accumulation and the sub-step count
	// (:814-820), the frame's ActiveEntry, and -- for an effect whose slot range is final and whose burst is small --
	// the rest of its bookkeeping. Effects that may have to grow, own more granules, read the slot mirror or take the
	// wide burst kernel are only noted here and finished in phase 2, in request order, on the calling thread.
	struct Slow { uint32_t entry; uint32_t req; HostEffect *E; HostEffectCold *C; };
	struct Part { std::vector<Slow> slow; uint64_t simSteps = 0; uint32_t maxSub = 0; };
	auto phase1 = [&](uint32_t r0, uint32_t r1, Part &part) {
		for (uint32_t r = r0; r < r1; r++) {
			ParticleSystem *sys = reqs[r].system;
			const SprFrameArgs &args = *reqs[r].args;
			const float clampedDt = args.dt < 0.1f ? args.dt : 0.1f; // sf::min(args.dt, 0.1f), :810
			ActiveEntry *out = entries.data() + work[r].firstActive;
			for (uint32_t i = 0; i < reqs[r].numActive; i++) {
				HostEffect &E = sys->effects[reqs[r].activeIds[i]];
				E.lastUpdateTime = args.gameTime;                // :814
				E.timeDelta += clampedDt;                        // :816
				uint32_t nsub = 0;
				while (E.timeDelta >= E.timeStep) { nsub++; E.timeDelta -= E.timeStep; } // :817-820
				out[i] = { E.global, nsub, 0, r << 2 }; // flags [31:2]: the effect's system in the work list
				if (!nsub) continue;
				if (nsub > part.maxSub) part.maxSub = nsub;
				const HostType &type = sys->types[E.typeId];
				if (!E.capacityFinal || (E.firstEmit && type.comp->burstAmount > 1024)) {
					part.slow.push_back({ work[r].firstActive + i, r, &E, &sys->effectsCold[reqs[r].activeIds[i]] });
					continue;
				}
				E.firstEmit = false;
				E.simSteps += nsub;
				part.simSteps += nsub;
				E.uploadFrameIndex = 0;                          // :679
				if (type.comp->localSpace) { HostEffectCold &C = sys->effectsCold[reqs[r].activeIds[i]]; C.boundsToWorld = C.particlesToWorld; } // gpuBounds uses the matrix of the simulated step (:675-677)
			}
		}
	};
	// (phase 1 on several host threads, by contiguous request ranges, was measured on the 100k-effect C5 frame: 0.90 ->
	// 0.75 ms with 4 threads, no gain in the step once the threads' start-up is paid -- the pass streams the effect
	// mirror through one core at memory speed, so the mirror was split into a 48-byte hot half instead)
	// Since the persistent workers of Context::hostPool the threads' start-up is no longer paid per frame: a frame of
	// >= 16384 listed effects runs this pass on up to four of them, by contiguous request ranges (phase 2 below walks the
	// parts in order, so everything that touches the context still happens in request order on the calling thread).
	const uint32_t passThreads = nested ? 1u : hostPassThreads(totalActive, numReqs);
	std::vector<Part> parts(passThreads);
	if (passThreads <= 1) phase1(0, numReqs, parts[0]);
	else {
		const std::vector<uint32_t> cut = splitRequests(reqs, numReqs, passThreads);
		hostPool.run(passThreads, [&](uint32_t t) { phase1(cut[t], cut[t + 1], parts[t]); });
	}

	auto tp1 = tNow();
	// Phase 2: everything that touches the context (ranges, granule ownership, the slot mirror, burst jobs), in request order
	for (Part &part : parts) {
		totalSimSteps += part.simSteps;
		if (part.maxSub > maxSub) maxSub = part.maxSub;
		for (const Slow &sl : part.slow) {
			ParticleSystem *sys = reqs[sl.req].system;
			HostEffect &E = *sl.E;
			HostEffectCold &C = *sl.C;
			ActiveEntry &ae = entries[sl.entry];
			const uint32_t nsub = ae.numSubsteps;
			const HostType &type = sys->types[E.typeId];
			const uint32_t burst = E.firstEmit ? type.comp->burstAmount : 0;
			// <= 20 timer spawns per sub-step (:478), each sub-step rounds up to one Particle4
			const uint64_t add = (uint64_t)nsub * (kMaxTimerSpawns + 3) + burst + 3;
			if (!E.capacityFinal) {
				ae.flags |= 2u; // k_finalize mirrors the exact slot count of this frame
				// re-base the bound on the newest exact count the device has mirrored (no synchronisation: an older
				// entry only means more allowances are added on top of it)
				const unsigned long long m = ((volatile const unsigned long long*)slotMirror)[E.global];
				const uint32_t mStamp = (uint32_t)(m >> 32);
				if (mStamp > C.createdStamp && mStamp >= C.evictedStamp) {
					uint64_t tight = (uint32_t)m;
					for (uint32_t k = 0; k < C.recentCount; k++) if (C.recent[k].stamp > mStamp) tight += C.recent[k].add;
					if (tight < C.slotUpperBound) C.slotUpperBound = (uint32_t)tight;
				}
				if (C.recentCount == 4) {
					C.evictedStamp = C.recent[0].stamp;
					C.recent[0] = C.recent[1]; C.recent[1] = C.recent[2]; C.recent[2] = C.recent[3];
					C.recentCount = 3;
				}
				C.recent[C.recentCount].stamp = thisStamp;
				C.recent[C.recentCount].add = add > 0xffffffffull ? 0xffffffffu : (uint32_t)add;
				C.recentCount++;
				const uint64_t need = (uint64_t)C.slotUpperBound + add;
				if (need > slotCapacityOf[E.global]) {
					if (E.simSteps == 0) { growEffect(E.global, pow2ceil(need), false); C.slotUpperBound = (uint32_t)need; ensureOwned(E.global, C.slotUpperBound); }
					else { resolve.push_back({ &E, &C, (uint32_t)add }); resolveGlobals.push_back(E.global); }
				} else {
					C.slotUpperBound = (uint32_t)need;
					ensureOwned(E.global, C.slotUpperBound);
				}
			}
			if (burst > 1024) {
				ae.flags |= 1u;
				uint32_t blocks = (burst + 255) / 256;
				burstJobs.push_back({ sl.entry, burstBlocks, blocks, 0 });
				burstBlocks += blocks;
			}
			E.firstEmit = false;
			E.simSteps += nsub;
			totalSimSteps += nsub;
			E.uploadFrameIndex = 0;                          // :679
			if (type.comp->localSpace) C.boundsToWorld = C.particlesToWorld; // gpuBounds uses the matrix of the simulated step (:675-677)
		}
	}
	// Phase 3: plan records are numbered in entry order
	for (ActiveEntry &ae : entries) { ae.planBase = numPlans; numPlans += ae.numSubsteps; }
	if (!resolve.empty()) {
		// the conservative bound ran out: read the exact slot counts back (rare) and grow if needed
		std::vector<RenderGather> got(resolve.size());
		gatherEffects(resolveGlobals.data(), (uint32_t)resolve.size(), got.data());
		for (size_t i = 0; i < resolve.size(); i++) {
			HostEffect &E = *resolve[i].E;
			HostEffectCold &C = *resolve[i].C;
			uint64_t need = (uint64_t)got[i].slotCount + resolve[i].add;
			if (need > hParams[E.global].slotCapacity) growEffect(E.global, pow2ceil(need * 2), true);
			C.slotUpperBound = (uint32_t)need;
			// the count just read is exact up to the last submitted frame: only this frame's allowance stays pending
			C.recent[0] = C.recent[C.recentCount - 1];
			C.recentCount = 1;
			C.evictedStamp = frameStamp;
			ensureOwned(E.global, C.slotUpperBound);
		}
	}

	auto t1 = tNow();
	// ---- sort-mode work lists
	std::vector<uint32_t> &segEffects = nested ? nestedSeg : scratchSegEffects;
	std::vector<SortTile> &sortTiles = nested ? nestedTiles : scratchSortTiles;
	segEffects.clear(); sortTiles.clear();
	uint32_t numBigSegments = 0, numBigTiles = 0, smallMaxSlots = 0;
	if (sortMode && numTinyEffects != numLiveEffects) { // (one-granule effects are sorted inside k_integrate_pass: a context of nothing else has no lists to build)
		// tile-path effects first, then the ones one CTA sorts whole (k_sort_small): ownedSlots is the effect's slot
		// bound rounded up to a tile, never below its real slot count
		static const bool smallSort = !(getenv("SPR_SMALL_SORT") && getenv("SPR_SMALL_SORT")[0] == '0');
		std::vector<uint32_t> &small = nested ? nestedSmall : scratchSmallEffects;
		small.clear();
		for (const ActiveEntry &ae : entries) {
			if (!ae.numSubsteps) continue;
			if (slotCapacityOf[ae.effect] == kGranuleSlots) continue; // sorted and expanded inside k_integrate_sort
			if (smallSort && ownedSlots[ae.effect] <= kSmallSortMax) { small.push_back(ae.effect); if (ownedSlots[ae.effect] > smallMaxSlots) smallMaxSlots = ownedSlots[ae.effect]; }
			else segEffects.push_back(ae.effect);
		}
		numBigSegments = (uint32_t)segEffects.size();
		segEffects.insert(segEffects.end(), small.begin(), small.end());
	}

	// ---- staging blob
	size_t bound = 256
		+ dirtyParams.size() * (sizeof(uint32_t) + sizeof(EffectParams))
		+ initStateIdx.size() * (sizeof(uint32_t) + sizeof(EffectState))
		+ typeIdx.size() * (sizeof(uint32_t) + sizeof(TypeDev))
		+ systemIdx.size() * (sizeof(uint32_t) + sizeof(SystemDev))
		+ ownerJobs.size() * sizeof(FillJob)
		+ entries.size() * sizeof(ActiveEntry) + work.size() * sizeof(SystemWork) + burstJobs.size() * sizeof(BurstJob)
		+ segEffects.size() * sizeof(uint32_t);
	if (sortMode) {
		for (uint32_t g : segEffects) bound += ((size_t)slotCapacityOf[g] / kSortTileKeys + 1) * sizeof(SortTile);
	}
	bound += 16 * 16;
	int par = stagingParity; stagingParity ^= 1;
	SPR_CUDA(cudaEventSynchronize(stagingEvent[par]));
	stagingH[par].reserve(bound);
	stagingD.reserve(bound, 0, stream);
	BlobWriter bw(stagingH[par].ptr);

	std::vector<EffectParams> &paramVals = nested ? nestedParamVals : scratchParamVals;
	paramVals.resize(dirtyParams.size());
	for (size_t i = 0; i < dirtyParams.size(); i++) { paramVals[i] = hParams[dirtyParams[i]]; dirtyParamsFlag[dirtyParams[i]] = 0; }
	size_t oParamIdx = bw.put(dirtyParams.data(), dirtyParams.size());
	size_t oParamVal = bw.put(paramVals.data(), paramVals.size());
	size_t oStateIdx = bw.put(initStateIdx.data(), initStateIdx.size());
	size_t oStateVal = bw.put(initStateVal.data(), initStateVal.size());
	size_t oTypeIdx = bw.put(typeIdx.data(), typeIdx.size());
	size_t oTypeVal = bw.put(typeVal.data(), typeVal.size());
	size_t oSysIdx = bw.put(systemIdx.data(), systemIdx.size());
	size_t oSysVal = bw.put(systemVal.data(), systemVal.size());
	// Fill jobs of one launch run in parallel; a range touched twice (freed, then re-owned) must see
	// its jobs in order, so the list is cut into batches without overlaps, launched one after another.
	std::vector<uint32_t> ownerBatchEnd;
	{
		std::map<uint32_t, uint32_t> inBatch; // start -> end of the ranges in the current batch
		for (size_t i = 0; i < ownerJobs.size(); i++) {
			uint32_t st = ownerJobs[i].start, en = st + ownerJobs[i].count;
			auto it = inBatch.upper_bound(st);
			bool overlap = (it != inBatch.end() && it->first < en);
			if (!overlap && it != inBatch.begin()) { --it; overlap = it->second > st; }
			if (overlap) { ownerBatchEnd.push_back((uint32_t)i); inBatch.clear(); }
			inBatch[st] = en;
		}
		ownerBatchEnd.push_back((uint32_t)ownerJobs.size());
	}
	size_t oOwner = bw.put(ownerJobs.data(), ownerJobs.size());
	size_t oEntries = bw.put(entries.data(), entries.size());
	size_t oWork = bw.put(work.data(), work.size());
	size_t oBurst = bw.put(burstJobs.data(), burstJobs.size());
	if (sortMode) {
		for (size_t s = 0; s < segEffects.size(); s++) {
			if (s == numBigSegments) numBigTiles = (uint32_t)sortTiles.size();
			// tiles up to the effect's slot bound (ownedSlots: never below its real slot count), not up to its
			// power-of-two capacity: C3's 10M slots are 2443 tiles of a 4096-tile range, and every empty CTA costs
			// each of the seven sort kernels a scheduling slot
			uint32_t tiles = (ownedSlots[segEffects[s]] + kSortTileKeys - 1) / kSortTileKeys;
			for (uint32_t t = 0; t < tiles; t++) sortTiles.push_back({ (uint32_t)s, t });
		}
		if (numBigSegments == segEffects.size()) numBigTiles = (uint32_t)sortTiles.size();
	}
	size_t oSeg = bw.put(segEffects.data(), segEffects.size());
	size_t oTiles = bw.put(sortTiles.data(), sortTiles.size());
	if (bw.size > bound) throw std::runtime_error("staging blob overflow");

	const uint32_t nParams = (uint32_t)dirtyParams.size(), nState = (uint32_t)initStateIdx.size();
	const uint32_t nTypes = (uint32_t)typeIdx.size(), nSys = (uint32_t)systemIdx.size(), nOwner = (uint32_t)ownerJobs.size();
	dirtyParams.clear(); ownerJobs.clear();
	clearUploads(initStateIdx, initStatePos); initStateVal.clear();
	clearUploads(typeIdx, typePos); typeVal.clear();
	clearUploads(systemIdx, systemPos); systemVal.clear();

	if (bw.size) {
		SPR_CUDA(cudaMemcpyAsync(stagingD.ptr, stagingH[par].ptr, bw.size, cudaMemcpyHostToDevice, stream));
		SPR_CUDA(cudaEventRecord(stagingEvent[par], stream));
	}
	uint8_t *D = stagingD.ptr;
#define DPTR(T, off) ((const T*)(D + (off)))
	if (nParams) { launch_scatter_words(stream, dParams.ptr, DPTR(uint32_t, oParamIdx), D + oParamVal, nParams, sizeof(EffectParams) / 4); launchCount++; }
	if (nState) { launch_scatter_words(stream, dState.ptr, DPTR(uint32_t, oStateIdx), D + oStateVal, nState, sizeof(EffectState) / 4); launchCount++; }
	if (nTypes) { launch_scatter_words(stream, dTypes.ptr, DPTR(uint32_t, oTypeIdx), D + oTypeVal, nTypes, sizeof(TypeDev) / 4); launchCount++; }
	if (nSys) { launch_scatter_words(stream, dSystems.ptr, DPTR(uint32_t, oSysIdx), D + oSysVal, nSys, sizeof(SystemDev) / 4); launchCount++; }
	for (uint32_t b0 = 0, bi = 0; nOwner && bi < ownerBatchEnd.size(); b0 = ownerBatchEnd[bi++]) {
		uint32_t cnt = ownerBatchEnd[bi] - b0;
		if (cnt) { launch_fill_jobs(stream, dGranuleOwner, DPTR(FillJob, oOwner) + b0, cnt); launchCount++; }
	}

	if (numPlans == 0) { SPR_CUDA(cudaGetLastError()); return; }

	auto t2 = tNow();
	// ---- the frame
	frameStamp++;
	dPlans.reserve(numPlans, 0, stream);
	PoolPtrs pool = poolPtrs();
	TablePtrs tab = tablePtrs();
	const ActiveEntry *dEntries = DPTR(ActiveEntry, oEntries);
	const uint32_t nEntries = (uint32_t)entries.size();
	SortScratch sortScratch = { nullptr, nullptr, nullptr, nullptr, deferExpansion, numSms, 0 };
	if (sortMode && !segEffects.empty()) {
		sortScratch.keysAlt = dKeys[1]; sortScratch.valsAlt = dVals[1];
		dSortHist.reserve(segEffects.size() * (4 * 256 + 1) + 4, 0, stream); // histograms, pass masks, four pass tickets
		dSortLookback.reserve(sortTiles.size() * 4 * 256 + 1, 0, stream);
		sortScratch.hist = dSortHist.ptr; sortScratch.lookback = dSortLookback.ptr;
		sortScratch.smallMaxSlots = smallMaxSlots;
		clear_sort_scratch(stream, sortScratch, (uint32_t)segEffects.size(), (uint32_t)sortTiles.size());
	}
	TIMED(kKPlan, launch_plan(stream, tab, DPTR(SystemWork, oWork), (uint32_t)work.size(), dEntries, frameStamp));
	// pool tiles up to the last owned granule, not up to the bump pointer: an effect's range is a power of two (C3: 10M
	// slots own 9767 tiles of a 16384-tile range) and the tiles behind its slot bound would only be looked at and dropped
	const uint64_t poolEnd = poolOwnedEnd < poolUsed ? poolOwnedEnd : poolUsed;
	const uint32_t poolTiles = (uint32_t)((poolEnd + kTileSlots - 1) / kTileSlots);
	for (uint32_t round = 0; round < maxSub; round++) {
		TIMED(kKEmit, launch_emit_warp(stream, pool, tab, DPTR(SystemWork, oWork), dEntries, nEntries, round, frameStamp));
		if (round == 0 && !burstJobs.empty()) {
			TIMED(kKEmitBurst, launch_emit_burst(stream, pool, tab, dEntries, DPTR(BurstJob, oBurst), (uint32_t)burstJobs.size(), burstBlocks));
		}
		epoch = (epoch + 1) & 0x3fffffffu;
		if (epoch == 0) { // status words carry a 30-bit epoch: wipe them when it wraps
			SPR_CUDA(cudaMemsetAsync(dTileStatus[0], 0, poolCapacity / kTileSlots * 8, stream));
			SPR_CUDA(cudaMemsetAsync(dTileStatus[1], 0, poolCapacity / kTileSlots * 8, stream));
			epoch = 1;
		}
		if (!sortMode && streamFused) {
			// stream mode, one pass: integrate + ranks + free list + 4x expansion in a software-pipelined persistent kernel
			TIMED(kKStreamFused, launch_stream_fused(stream, pool, tab, poolTiles, frameStamp, round, epoch, typesWithGravityPoints != 0, numSms, streamFusedCtasPerSm, ticketBase[0], vertexMapValid ? &vertexMap : nullptr));
		} else {
			TIMED(kKIntegrate, launch_integrate(stream, pool, tab, poolTiles, frameStamp, round, epoch, sortMode, typesWithGravityPoints != 0, numTinyEffects != 0));
			TIMED(kKDeadAppend, launch_dead_append(stream, pool, tab, poolTiles, frameStamp, round, epoch, !sortMode, ticketBase[1]));
		}
	}
	if (!sortMode && !streamFused) TIMED(kKExpandStream, launch_expand_stream(stream, pool, tab, poolTiles, frameStamp));
	TIMED(kKFinalize, launch_finalize(stream, tab, dEntries, nEntries, frameStamp));

	if (sortMode && !segEffects.empty()) {
		const SortScratch &sc = sortScratch;
		SortTimingHooks hooks = { this, [](void *c, int k) -> void* { return (void*)((Context*)c)->timingBegin(k == 4 ? kKSortSmall : kKSortBase + k); },
			[](void *c, int k, void *b) { ((Context*)c)->timingEnd(k == 4 ? kKSortSmall : kKSortBase + k, (cudaEvent_t)b); } };
		launchCount += launch_sort_and_expand(stream, pool, tab, sc, DPTR(uint32_t, oSeg), (uint32_t)segEffects.size(), numBigSegments,
			DPTR(SortTile, oTiles), (uint32_t)sortTiles.size(), numBigTiles, timingEnabled ? &hooks : nullptr);
	}
#undef DPTR
	if (hostTiming) {
		auto t3 = tNow();
		auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count() * 1e-3; };
		static double acc[5]; static int frames;
		acc[0] += us(t0, t1); acc[1] += us(t1, t2); acc[2] += us(t2, t3); acc[3] += us(tv, t0); acc[4] += us(t0, tp1);
		if (++frames % 16 == 0) {
			fprintf(stderr, "[spr host] per frame: id check %.0f us, effect pass %.0f us (phase 1: %.0f us), work lists + staging %.0f us, launches %.0f us\n",
				acc[3] / 16, acc[0] / 16, acc[4] / 16, acc[1] / 16, acc[2] / 16);
			for (double &a : acc) a = 0;
		}
	}
	SPR_CUDA(cudaGetLastError());
}

uint32_t *Context::gatherIndexBuffer(uint32_t n)
{
	gatherIdxH.reserve(n ? n : 1);
	return gatherIdxH.ptr;
}

// Counts + bounds of the n effects whose context-wide indices the caller wrote into gatherIndexBuffer(n): one H2D
// of the indices, one kernel, one D2H, all through pinned memory. The result stays valid until the next gather.
const RenderGather *Context::gatherSubmit(uint32_t n)
{
	if (!n) return gatherH.ptr;
	SPR_CUDA(cudaSetDevice(device));
	if (!dirtyParams.empty() || !initStateIdx.empty() || !typeIdx.empty() || !systemIdx.empty() || !ownerJobs.empty()) flushUploads();
	gatherIdxD.reserve(n, 0, stream);
	gatherD.reserve(n, 0, stream);
	gatherH.reserve(n);
	SPR_CUDA(cudaMemcpyAsync(gatherIdxD.ptr, gatherIdxH.ptr, n * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
	launch_gather_render(stream, tablePtrs(), gatherIdxD.ptr, n, gatherD.ptr); launchCount++;
	SPR_CUDA(cudaMemcpyAsync(gatherH.ptr, gatherD.ptr, n * sizeof(RenderGather), cudaMemcpyDeviceToHost, stream));
	SPR_CUDA(cudaStreamSynchronize(stream));
	return gatherH.ptr;
}

void Context::gatherEffects(const uint32_t *globals, uint32_t n, RenderGather *out)
{
	if (!n) return;
	memcpy(gatherIndexBuffer(n), globals, n * sizeof(uint32_t));
	memcpy(out, gatherSubmit(n), n * sizeof(RenderGather));
}

// ---------------------------------------------------------------------------
// ParticleSystem
// ---------------------------------------------------------------------------

ParticleSystem::ParticleSystem(Context *c, const SprSystemsDesc &desc)
	: ctx(c), initRng(1, 1), bufferFrameIndex(c->maxParticleCacheFrames)
{
	if (!ctx->freeSystemGlobals.empty()) { systemIdx = ctx->freeSystemGlobals.back(); ctx->freeSystemGlobals.pop_back(); }
	else {
		systemIdx = ctx->numSystemGlobals++;
		ctx->dSystems.reserve(ctx->numSystemGlobals, ctx->numSystemGlobals - 1, ctx->stream);
	}
	SystemDev sd;
	sd.rngState = desc.seed[1];            // sf::Random(desc.seed[1], 581271), :267; ctor src/sf/Random.h:11
	sd.rngInc = 581271ull | 1ull;
	Context::queueUpload(ctx->systemIdx, ctx->systemVal, ctx->systemPos, systemIdx, sd);
	ctx->systemsAlive.push_back(this);
}

ParticleSystem::~ParticleSystem()
{
	for (uint32_t id = 0; id < effects.size(); id++) if (effects[id].inUse) remove(id);
	for (HostType &t : types) if (t.comp) { if (t.hasGravityPoints) ctx->typesWithGravityPoints--; ctx->freeTypeGlobals.push_back(t.global); t.comp = nullptr; }
	ctx->freeSystemGlobals.push_back(systemIdx);
	ctx->systemsAlive.erase(std::find(ctx->systemsAlive.begin(), ctx->systemsAlive.end(), this));
}

uint32_t ParticleSystem::reserveEffectType(const SprParticleSystemComponent *c)
{
	uint32_t typeId;
	auto it = componentToType.find((const void*)c);
	if (it == componentToType.end()) {
		if (!freeTypeIds.empty()) { typeId = freeTypeIds.back(); freeTypeIds.pop_back(); }
		else { typeId = (uint32_t)types.size(); types.emplace_back(); }
		componentToType[(const void*)c] = typeId;
		HostType &type = types[typeId];
		type.comp = c;
		type.refCount = 0;
		type.renderOrder = c->renderOrder;
		type.tiebreaker = typeId;
		type.timeStep = c->timeStep;

		// initEffectType (:298-450): LUT + type UBO on the host, sampler constants for the device
		bakeTypeUbo(type.typeUbo, typeId, *c);
		memset(type.lut, 0, sizeof(type.lut));
		bakeSplineRow(type.lut, 0, c->scaleSpline, false);
		bakeSplineRow(type.lut, 1, c->alphaSpline, false);
		bakeSplineRow(type.lut, 2, c->additiveSpline, true);
		bakeSplineRow(type.lut, 3, c->erosionSpline, false);
		bakeGradientRow(type.lut + SPR_SPLINE_SAMPLE_RATE * 4, c->gradient);

		if (!ctx->freeTypeGlobals.empty()) { type.global = ctx->freeTypeGlobals.back(); ctx->freeTypeGlobals.pop_back(); }
		else {
			type.global = ctx->numTypeGlobals++;
			ctx->dTypes.reserve(ctx->numTypeGlobals, ctx->numTypeGlobals - 1, ctx->stream);
		}
		TypeDev td;
		memset(&td, 0, sizeof(td));
		bakeRandomVec3(td.emitPosition, c->emitPosition);
		bakeRandomVec3(td.emitVelocity, c->emitVelocity);
		td.attractorOffset[0] = c->emitVelocityAttractorOffset.x;
		td.attractorOffset[1] = c->emitVelocityAttractorOffset.y;
		td.attractorOffset[2] = c->emitVelocityAttractorOffset.z;
		td.attractorStrength = c->emitVelocityAttractorStrength;
		td.spawnTime = c->spawnTime;
		td.spawnTimeVariance = c->spawnTimeVariance;
		td.burstAmount = c->burstAmount;
		// draws of one timer spawn: timer + position + velocity + seed (:484, :498-499, :522)
		td.drawsPerSpawn = 1 + drawsOfSampler(td.emitPosition) + drawsOfSampler(td.emitVelocity) + 1;
		lcgJump(td.drawsPerSpawn - 1, td.mulKm1, td.sumKm1);
		for (uint32_t j = 0; j <= kMaxTimerSpawns; j++) lcgJump((uint64_t)j * td.drawsPerSpawn, td.mulJK[j], td.sumJK[j]);
		td.drag = c->drag;
		td.lifeTime = c->lifeTime;
		td.lifeTimeVariance = c->lifeTimeVariance * (1.0f / 16777216.0f);   // :535
		td.cullPadding = c->cullPadding;
		td.localSpace = c->localSpace ? 1u : 0u;
		td.numGravityPoints = c->numGravityPoints < SPR_MAX_GRAVITY_POINTS ? c->numGravityPoints : SPR_MAX_GRAVITY_POINTS;
		for (uint32_t i = 0; i < td.numGravityPoints; i++) {               // :540-550
			const SprGravityPoint &p = c->gravityPoints[i];
			td.gravityPoints[i].position[0] = p.position.x;
			td.gravityPoints[i].position[1] = p.position.y;
			td.gravityPoints[i].position[2] = p.position.z;
			td.gravityPoints[i].radius = p.radius > 0.05f ? p.radius : 0.05f; // sf::max(p.radius, 0.05f)
			td.gravityPoints[i].strength = -p.strength;
		}
		type.hasGravityPoints = td.numGravityPoints != 0;
		if (type.hasGravityPoints) ctx->typesWithGravityPoints++;
		Context::queueUpload(ctx->typeIdx, ctx->typeVal, ctx->typePos, type.global, td);
	} else {
		typeId = it->second;
	}
	types[typeId].refCount++;
	return typeId;
}

void ParticleSystem::releaseEffectTypeImp(uint32_t typeId)
{
	HostType &type = types[typeId];
	if (--type.refCount == 0) {
		componentToType.erase((const void*)type.comp);
		if (type.hasGravityPoints) ctx->typesWithGravityPoints--;
		ctx->freeTypeGlobals.push_back(type.global);
		type = HostType();
		freeTypeIds.push_back(typeId);
	}
}

void ParticleSystem::releaseEffectType(const SprParticleSystemComponent *c)
{
	auto it = componentToType.find((const void*)c);
	if (it != componentToType.end()) releaseEffectTypeImp(it->second);
}

uint32_t ParticleSystem::addEffect(uint32_t entityId, uint8_t componentIndex, const SprParticleSystemComponent *c, const SprTransform &transform)
{
	(void)entityId; (void)componentIndex;
	uint32_t effectId;
	if (!freeEffectIds.empty()) { effectId = freeEffectIds.back(); freeEffectIds.pop_back(); }
	else { effectId = (uint32_t)effects.size(); effects.emplace_back(); effectsCold.emplace_back(); activeMark.push_back(kMarkFree); }
	activeMark[effectId] = 0;

	uint32_t typeId = reserveEffectType(c);
	const HostType &type = types[typeId];

	HostEffect &E = effects[effectId];
	HostEffectCold &C = effectsCold[effectId];
	E = HostEffect();
	C = HostEffectCold();
	E.inUse = true;
	E.typeId = typeId;
	E.timeStep = type.timeStep * (1.0f + initRng.nextFloat() * 0.1f);   // :744
	C.origin = transform.position;
	E.instantDelete = c->instantDelete != 0;
	C.particlesToWorld = identity34();
	C.boundsToWorld = identity34();
	E.uploadFrameIndex = ctx->maxParticleCacheFrames;
	E.global = ctx->allocEffectGlobal();
	C.createdStamp = ctx->frameStamp; // mirror entries up to this frame belong to whoever had the index before

	EffectParams &P = ctx->hParams[E.global];
	memset(&P, 0, sizeof(P));
	P.typeIdx = type.global;
	P.systemIdx = systemIdx;
	P.timeStep = E.timeStep;
	P.gravity[0] = c->gravity.x; P.gravity[1] = c->gravity.y; P.gravity[2] = c->gravity.z;
	P.effectId = effectId;
	P.live = 1;
	// the sphere area the reference registers with the area system (:762-763)
	P.areaSphere[0] = transform.position.x; P.areaSphere[1] = transform.position.y; P.areaSphere[2] = transform.position.z;
	P.areaSphere[3] = c->updateRadius;
	SprMat34 entityToWorld = transformToMatrix(transform);
	if (c->localSpace) {                                                  // :755-760
		writeColMajor44(identity34(), P.emitterToWorld);
		C.particlesToWorld = entityToWorld;
	} else {
		writeColMajor44(entityToWorld, P.emitterToWorld);
	}
	P.slotBase = 0; P.slotCapacity = 0;
	ctx->slotCapacityOf[E.global] = 0;
	ctx->growEffect(E.global, pow2ceil((uint64_t)c->burstAmount + 64), false);
	{
		// A slot is held from the sub-step that spawns its particle until the integrate pass after the one in which its
		// life reaches 0 (:556-567): life = 1 falls by timeStep / L per sub-step with L = lifeTime + seed * variance * 2^-24
		// <= lifeTime + variance (:535, :596-605), and an effect's timeStep is never below its type's (:744). With at most
		// 20 timer spawns per sub-step (:478) plus the one burst, the slot count can never exceed
		//   burst + 20 * (floor(Lmax / typeTimeStep) + 4) + 4
		// (two sub-steps of slack for rounding in the life accumulation, two for the spawn / free order inside a
		// sub-step, +4 for the Particle4 granularity). When that fits the range just allocated, the per-frame capacity
		// bookkeeping of Context::update (mirror read, allowances, growth checks) is skipped for this effect for good.
		const double lmax = (double)c->lifeTime + (double)c->lifeTimeVariance;
		if (c->lifeTime > 0.0f && c->lifeTimeVariance >= 0.0f && type.timeStep > 0.0f && std::isfinite(lmax)) {
			const double steps = std::floor(lmax / (double)type.timeStep) + 4.0;
			const double bound = (double)c->burstAmount + 20.0 * steps + 4.0;
			// (only where owning the range up to the bound from the start costs at most one tile of slots the effect may
			// never reach: the streaming kernels load every owned granule)
			const uint32_t cap = ctx->slotCapacityOf[E.global];
			if (bound <= (double)cap && (cap <= kTileSlots || 20.0 * steps + 8.0 <= (double)kTileSlots)) {
				E.capacityFinal = true;
				C.slotUpperBound = (uint32_t)bound;
				ctx->ensureOwned(E.global, C.slotUpperBound);
			}
		}
	}

	EffectState S;
	memset(&S, 0, sizeof(S));
	S.firstEmit = 1;
	S.emitOnTimer = c->emitterOnTime;                                      // :750
	Context::queueUpload(ctx->initStateIdx, ctx->initStateVal, ctx->initStatePos, E.global, S);
	ctx->markParamsDirty(E.global);
	return effectId;
}

void ParticleSystem::updateTransform(uint32_t effectId, const SprTransformUpdate &update)
{
	HostEffect &E = effects[effectId];
	HostEffectCold &C = effectsCold[effectId];
	const HostType &type = types[E.typeId];
	EffectParams &P = ctx->hParams[E.global];
	if (type.comp->localSpace) {
		C.particlesToWorld = update.entityToWorld;
	} else {
		writeColMajor44(update.entityToWorld, P.emitterToWorld);
	}
	C.origin = update.transform.position;
	// updateSphereArea(areaId, { position, type.updateRadius }), :782-783
	P.areaSphere[0] = update.transform.position.x; P.areaSphere[1] = update.transform.position.y; P.areaSphere[2] = update.transform.position.z;
	P.areaSphere[3] = type.comp->updateRadius;
	ctx->markParamsDirty(E.global);
}

bool ParticleSystem::prepareForRemove(uint32_t effectId, const SprFrameArgs &args)
{
	HostEffect &E = effects[effectId];
	if (!E.hostStopEmit) {
		E.hostStopEmit = true;                                             // :791
		ctx->hParams[E.global].hostStopEmit = 1;
		ctx->markParamsDirty(E.global);
	}
	if (E.instantDelete || args.gameTime - E.lastUpdateTime > 2.0) return true; // :793 (order-free ||)
	RenderGather g;
	ctx->gatherEffects(&E.global, 1, &g);
	return g.aliveCount == 0;
}

void ParticleSystem::remove(uint32_t effectId)
{
	HostEffect &E = effects[effectId];
	if (!E.inUse) return;
	releaseEffectTypeImp(E.typeId);
	EffectParams &P = ctx->hParams[E.global];
	if (P.slotCapacity) ctx->freeRange(P.slotBase, P.slotCapacity);
	if (P.slotCapacity == kGranuleSlots) ctx->numTinyEffects--;
	P.slotCapacity = 0;
	P.live = 0;                       // removeSphereArea, :801
	ctx->markParamsDirty(E.global);
	ctx->slotCapacityOf[E.global] = 0;
	ctx->ownedSlots[E.global] = 0;
	ctx->freeEffectGlobals.push_back(E.global);
	ctx->effectLive[E.global] = 0;
	ctx->numLiveEffects--;
	E = HostEffect();
	activeMark[effectId] = kMarkFree;
	freeEffectIds.push_back(effectId);
}

void ParticleSystem::updateParticles(const uint32_t *activeIds, uint32_t numActive, const SprFrameArgs &args)
{
	UpdateRequest req = { this, activeIds, numActive, &args };
	ctx->update(&req, 1);
}

SprBounds3 boundsFromGather(const RenderGather &g, const HostType &type, const HostEffectCold &C)
{
	// :672-677
	SprBounds3 b;
	b.origin.x = (g.boundsMin[0] + g.boundsMax[0]) * 0.5f;
	b.origin.y = (g.boundsMin[1] + g.boundsMax[1]) * 0.5f;
	b.origin.z = (g.boundsMin[2] + g.boundsMax[2]) * 0.5f;
	b.extent.x = (g.boundsMax[0] - g.boundsMin[0]) * 0.5f + type.comp->cullPadding;
	b.extent.y = (g.boundsMax[1] - g.boundsMin[1]) * 0.5f + type.comp->cullPadding;
	b.extent.z = (g.boundsMax[2] - g.boundsMin[2]) * 0.5f + type.comp->cullPadding;
	if (type.comp->localSpace) b = transformBounds(C.boundsToWorld, b);
	return b;
}

void ParticleSystem::renderMain(const uint32_t *visibleIds, uint32_t numVisible, const SprRenderArgs &args, SprRenderOutput &out)
{
	uint32_t *globals = ctx->gatherIndexBuffer(numVisible);
	for (uint32_t i = 0; i < numVisible; i++) globals[i] = effects[visibleIds[i]].global;
	renderMainWith(visibleIds, numVisible, args, ctx->gatherSubmit(numVisible), out);
}

// renderMain for many systems with ONE device read-back (counts + bounds of every visible effect); the indices go
// out of and the records come back into pinned memory that the per-system passes read in place.
void Context::renderBatch(ParticleSystem *const *systems, uint32_t numSystems, const uint32_t *const *visibleIds,
	const uint32_t *numVisible, const SprRenderArgs *args, size_t argsStride, SprRenderOutput *outs)
{
	uint64_t total = 0;
	for (uint32_t s = 0; s < numSystems; s++) total += numVisible[s];
	if (total > 0xffffffffull) throw std::runtime_error("renderBatch: more than 2^32 visible effects");
	uint32_t *globals = gatherIndexBuffer((uint32_t)total);
	size_t n = 0;
	for (uint32_t s = 0; s < numSystems; s++) {
		const std::vector<HostEffect> &effects = systems[s]->effects;
		const uint32_t *ids = visibleIds[s];
		for (uint32_t i = 0; i < numVisible[s]; i++) globals[n++] = effects[ids[i]].global;
	}
	const RenderGather *got = gatherSubmit((uint32_t)total);
	// The per-system passes (effect sort, draw records) touch nothing but their own system: a frame with tens of
	// thousands of visible effects in many systems (C5: 100k effects, 20 MB of draw records) is split over a few host
	// threads, 16 systems at a time. SPR_RENDER_THREADS=1 keeps everything on the calling thread.
	const uint32_t maxThreads = HostPool::maxThreads();
	const uint32_t kChunk = 16;
	uint32_t nt = (total >= 32768 && numSystems >= 4 * kChunk) ? maxThreads : 1;
	if (nt > (numSystems + kChunk - 1) / kChunk) nt = (numSystems + kChunk - 1) / kChunk;
	if (nt > 1) {
		// a system named twice would be written by two threads: such a batch stays on the calling thread
		const uint64_t stamp = ++renderBatchStamp;
		for (uint32_t s = 0; s < numSystems; s++) {
			if (systems[s]->renderBatchStamp == stamp) { nt = 1; break; }
			systems[s]->renderBatchStamp = stamp;
		}
	}
	if (nt <= 1) {
		size_t off = 0;
		for (uint32_t s = 0; s < numSystems; s++) {
			const SprRenderArgs &a = *(const SprRenderArgs*)((const char*)args + argsStride * s);
			systems[s]->renderMainWith(visibleIds[s], numVisible[s], a, got + off, outs[s]);
			off += numVisible[s];
		}
		return;
	}
	std::vector<size_t> offs(numSystems);
	{ size_t off = 0; for (uint32_t s = 0; s < numSystems; s++) { offs[s] = off; off += numVisible[s]; } }
	std::atomic<uint32_t> next(0);
	hostPool.run(nt, [&](uint32_t) {
		for (;;) {
			const uint32_t s0 = next.fetch_add(kChunk);
			if (s0 >= numSystems) break;
			const uint32_t s1 = s0 + kChunk < numSystems ? s0 + kChunk : numSystems;
			for (uint32_t s = s0; s < s1; s++) {
				const SprRenderArgs &a = *(const SprRenderArgs*)((const char*)args + argsStride * s);
				systems[s]->renderMainWith(visibleIds[s], numVisible[s], a, got + offs[s], outs[s]);
			}
		}
	});
}

void ParticleSystem::renderMainWith(const uint32_t *visibleIds, uint32_t numVisible, const SprRenderArgs &args, const RenderGather *got, SprRenderOutput &out)
{
	uint32_t frameParticlesLeft = ctx->maxParticlesPerFrame;              // :827
	bufferFrameIndex++;                                                   // :831
	const uint64_t frame = bufferFrameIndex;

	std::vector<SortedEffect> &sorted = sortedScratch;
	sorted.clear();
	sorted.reserve(numVisible);
	for (uint32_t i = 0; i < numVisible; i++) {                           // :835-846
		uint32_t effectId = visibleIds[i];
		const HostEffect &E = effects[effectId];
		const HostType &type = types[E.typeId];
		uint32_t numParticles = got[i].aliveCount;
		if (numParticles == 0) continue;
		if (!frustumIntersects(args.frustum, boundsFromGather(got[i], type, effectsCold[effectId]))) continue;
		SortedEffect s;
		s.renderOrder = type.renderOrder;
		s.tiebreaker = type.tiebreaker;
		const HostEffectCold &C = effectsCold[effectId];
		float dx = args.cameraPosition.x - C.origin.x, dy = args.cameraPosition.y - C.origin.y, dz = args.cameraPosition.z - C.origin.z;
		s.depth = dx*dx + dy*dy + dz*dz;                                  // sf::lengthSq, :844
		s.effectId = effectId;
		s.numParticles = numParticles;
		sorted.push_back(s);
	}
	std::sort(sorted.begin(), sorted.end());                              // :848 (strict total order)

	drawRecords.clear();
	drawRecords.resize(sorted.size());                                    // value-initialised: every record starts zeroed
	uint32_t prevTypeId = ~0u, numRecords = 0;
	const float aspect = args.viewToClip.v[0] / args.viewToClip.v[5];
	for (const SortedEffect &s : sorted) {                                // :851-902
		HostEffect &E = effects[s.effectId];
		uint32_t numParticles = s.numParticles;
		bool uploaded = false;
		if (frame - E.uploadFrameIndex >= ctx->maxParticleCacheFrames) {  // :861-867
			if (numParticles >= frameParticlesLeft) break;               // `return`, :862
			frameParticlesLeft -= numParticles;
			E.uploadFrameIndex = frame;
			uploaded = true;
		}
		SprDrawRecord &r = drawRecords[numRecords++];                     // filled in place
		r.uploadedThisFrame = uploaded ? 1u : 0u;
		r.uploadRingSlot = (uint32_t)(E.uploadFrameIndex % ctx->maxParticleCacheFrames); // :869
		r.effectId = s.effectId;
		r.typeId = E.typeId;
		r.numParticles = numParticles;
		r.vertexByteOffset = (uint64_t)ctx->hParams[E.global].slotBase * 128ull;
		writeColMajor44(effectsCold[s.effectId].particlesToWorld, r.instanceUbo.u_ParticleToWorld); // :871-876
		memcpy(r.instanceUbo.u_WorldToClip, args.worldToClip.v, sizeof(float) * 16);
		r.instanceUbo.u_Aspect = aspect;
		r.instanceUbo.u_InvDelta = E.timeStep - E.timeDelta;
		r.instanceUbo.u_VisFogMad = args.visFogWorldMad;
		if (E.typeId != prevTypeId) { r.typeUboChanged = 1; prevTypeId = E.typeId; } // :880-883
	}
	drawRecords.resize(numRecords);
	out.deviceVertices = (const SprGpuParticle*)ctx->dVertices;
	out.records = drawRecords.data();
	out.numRecords = (uint32_t)drawRecords.size();
	out.numSortedEffects = (uint32_t)sorted.size();
}

// ---------------------------------------------------------------------------
// Vertex stage
// ---------------------------------------------------------------------------

uint64_t ParticleSystem::shadeVertices(const SprDrawRecord *records, uint32_t numRecords, const SprFogImage *fog, SprShadedVertex *deviceOut, uint64_t capacityVertices)
{
	static_assert(sizeof(SprVertexInstanceUbo) == 160 && sizeof(SprVertexTypeUbo) == 96, "UBO layouts of src/game/shader/Particle.h:100-125");
	static_assert(sizeof(SprShadedVertex) == 64, "four float4 per shaded vertex");
	if (!numRecords) return 0;
	SPR_CUDA(cudaSetDevice(ctx->device));
	std::vector<ShadeJob> jobs;
	std::vector<uint32_t> luts;                         // 128 words per distinct effect type
	std::unordered_map<uint32_t, uint32_t> lutOfType;
	uint64_t vertex = 0;
	uint32_t blocks = 0;
	for (uint32_t r = 0; r < numRecords; r++) {
		const SprDrawRecord &d = records[r];
		if (d.typeId >= types.size() || !types[d.typeId].comp) throw std::runtime_error("sprShadeVertices: record with an unknown effect type");
		uint64_t n = d.numParticles;
		if (vertex + 4 * n > capacityVertices) n = (capacityVertices - vertex) / 4;
		if (!n) break;
		auto it = lutOfType.find(d.typeId);
		if (it == lutOfType.end()) {
			it = lutOfType.emplace(d.typeId, (uint32_t)(luts.size() / 128)).first;
			const uint8_t *lut = types[d.typeId].lut;
			for (uint32_t i = 0; i < 128; i++) luts.push_back((uint32_t)lut[4 * i] | (uint32_t)lut[4 * i + 1] << 8 | (uint32_t)lut[4 * i + 2] << 16 | (uint32_t)lut[4 * i + 3] << 24);
		}
		ShadeJob j;
		memset(&j, 0, sizeof(j));
		j.srcRecord = d.vertexByteOffset / sizeof(SprGpuParticle);
		j.dstVertex = vertex;
		j.numParticles = (uint32_t)n;
		j.lutIndex = it->second;
		j.firstBlock = blocks;
		memcpy(j.instance, &d.instanceUbo, sizeof(j.instance));
		memcpy(j.type, &types[d.typeId].typeUbo, sizeof(j.type));
		jobs.push_back(j);
		blocks += (uint32_t)((n + 255) / 256);
		vertex += 4 * n;
	}
	if (jobs.empty()) return 0;
	const size_t jobBytes = jobs.size() * sizeof(ShadeJob), lutBytes = luts.size() * sizeof(uint32_t);
	ctx->shadeBlobD.reserve(jobBytes + lutBytes, 0, ctx->stream);
	// pageable sources: the copies are staged by the runtime before the calls return
	SPR_CUDA(cudaMemcpyAsync(ctx->shadeBlobD.ptr, jobs.data(), jobBytes, cudaMemcpyHostToDevice, ctx->stream));
	SPR_CUDA(cudaMemcpyAsync(ctx->shadeBlobD.ptr + jobBytes, luts.data(), lutBytes, cudaMemcpyHostToDevice, ctx->stream));
	launch_shade_vertices(ctx->stream, (const float4*)ctx->dVertices, (const ShadeJob*)ctx->shadeBlobD.ptr, (uint32_t)jobs.size(), blocks,
		(const uint32_t*)(ctx->shadeBlobD.ptr + jobBytes), fog ? (const uint8_t*)fog->deviceTexels : nullptr, fog ? fog->width : 0, fog ? fog->height : 0,
		(float4*)deviceOut, capacityVertices);
	ctx->launchCount++;
	SPR_CUDA(cudaGetLastError());
	return vertex;
}

// ---------------------------------------------------------------------------
// BillboardSystem
// ---------------------------------------------------------------------------

BillboardSystem::BillboardSystem(Context *c, uint32_t maxBillboardsPerFrame)
	: ctx(c), maxPerFrame(maxBillboardsPerFrame ? maxBillboardsPerFrame : 4096)
{
}

// packChannel / packColor, BillboardSystem.cpp:12-28: sf::clamp((uint32_t)(f * 255.0f), 0u, 255u). The
// float -> uint32_t conversion of a negative or huge value is undefined in C++; on x86-64 it is the low 32
// bits of the 64-bit truncation (cvttss2si), reproduced here so that out-of-range colours pack alike.
static uint32_t packChannel(float f)
{
	float v = f * 255.0f;
	uint32_t u;
	if (v >= -9223372036854775808.0f && v < 9223372036854775808.0f) u = (uint32_t)(uint64_t)(int64_t)v;
	else u = 0; // cvttss2si's "integer indefinite" 0x8000000000000000 -> low word 0 (also NaN)
	return u > 255u ? 255u : u;
}

static_assert(sizeof(SprBillboardPacked) == sizeof(BillboardDev) && offsetof(SprBillboardPacked, transform) == offsetof(BillboardDev, transform)
	&& offsetof(SprBillboardPacked, color) == offsetof(BillboardDev, color) && offsetof(SprBillboardPacked, depth) == offsetof(BillboardDev, depth),
	"SprBillboardPacked is BillboardDev");

// addBillboard's conversion (:80-117) of billboards [i0, i1) into packed form; returns the end of what it wrote
BillboardDev *BillboardSystem::pack(const SprBillboard *billboards, uint32_t i0, uint32_t i1, BillboardDev *dst)
{
	for (uint32_t i = i0; i < i1; i++) {
		const SprBillboard &b = billboards[i];
		if (!b.sprite.atlas) continue;                              // !sprite || !sprite->isLoaded(), :82, :100
		BillboardDev &d = *dst++;
		d.atlas = b.sprite.atlas;
		d.minVert[0] = b.sprite.minVert.x; d.minVert[1] = b.sprite.minVert.y;
		d.maxVert[0] = b.sprite.maxVert.x; d.maxVert[1] = b.sprite.maxVert.y;
		d.vertUvScale[0] = b.sprite.vertUvScale.x; d.vertUvScale[1] = b.sprite.vertUvScale.y;
		d.vertUvBias[0] = b.sprite.vertUvBias.x; d.vertUvBias[1] = b.sprite.vertUvBias.y;
		memcpy(d.transform, b.transform, sizeof(d.transform));
		d.anchor[0] = b.anchor.x; d.anchor[1] = b.anchor.y;
		d.color = packChannel(b.color.x) | packChannel(b.color.y) << 8 | packChannel(b.color.z) << 16 | packChannel(b.color.w) << 24; // :22-28
		d.cropMin[0] = b.cropMin.x; d.cropMin[1] = b.cropMin.y;
		d.cropMax[0] = b.cropMax.x; d.cropMax[1] = b.cropMax.y;
		d.depth = b.depth;                                          // order.index = billboards.size - 1 = its position in the frame, :95, :115
	}
	return dst;
}

void BillboardSystem::addPacked(const BillboardDev *packed, uint32_t count, bool onDevice)
{
	if (!count) return;
	segments.push_back({ onDevice ? 2 : 1, packed, 0, count });
	frameCount += count;
}

void BillboardSystem::add(const SprBillboard *billboards, uint32_t count)
{
	if (pendingCount + count > pendingCapacity) {
		// the frame's billboards are staged in pinned memory (one DMA in renderMain); nothing is in flight between frames
		size_t cap = pendingCapacity ? pendingCapacity : 4096;
		while (cap < pendingCount + count) cap *= 2;
		BillboardDev *p = nullptr;
		SPR_CUDA(cudaSetDevice(ctx->device));
		SPR_CUDA(cudaMallocHost(&p, cap * sizeof(BillboardDev)));
		if (pendingCount) memcpy(p, pending, pendingCount * sizeof(BillboardDev));
		if (pending) cudaFreeHost(pending);
		pending = p; pendingCapacity = cap;
	}
	auto convert = [&](uint32_t i0, uint32_t i1, BillboardDev *dst) { return pack(billboards, i0, i1, dst); };
	const size_t before = pendingCount;
	auto noteSegment = [&] { // the converted run joins the frame's list (merged with a converted run right before it)
		const size_t added = pendingCount - before;
		if (!added) return;
		if (!segments.empty() && segments.back().kind == 0) segments.back().count += added;
		else segments.push_back({ 0, nullptr, before, added });
		frameCount += added;
	};
	// A large call (the 1M-billboard frame: 220 MB through the host) is converted by a few threads: each counts the
	// loaded sprites of its range, the counts give every range its place in the staging array, then the ranges are
	// converted side by side -- the order of the billboards is the order of the call either way.
	const uint32_t nt = count >= 65536 ? HostPool::maxThreads() : 1;
	if (nt <= 1) {
		pendingCount = (size_t)(convert(0, count, pending + pendingCount) - pending);
		noteSegment();
		return;
	}
	std::vector<uint32_t> cut(nt + 1), valid(nt, 0);
	for (uint32_t t = 0; t <= nt; t++) cut[t] = (uint32_t)((uint64_t)count * t / nt);
	auto run = [&](auto &&fn) { ctx->hostPool.run(nt, fn); };
	run([&](uint32_t t) { uint32_t v = 0; for (uint32_t i = cut[t]; i < cut[t + 1]; i++) v += billboards[i].sprite.atlas != 0; valid[t] = v; });
	std::vector<size_t> first(nt + 1, pendingCount);
	for (uint32_t t = 0; t < nt; t++) first[t + 1] = first[t] + valid[t];
	run([&](uint32_t t) { convert(cut[t], cut[t + 1], pending + first[t]); });
	pendingCount = first[nt];
	noteSegment();
}

void BillboardSystem::renderMain(const SprRenderArgs &args, SprBillboardOutput &out)
{
	SPR_CUDA(cudaSetDevice(ctx->device));
	memset(&out, 0, sizeof(out));
	memcpy(out.worldToClip, args.worldToClip.v, sizeof(out.worldToClip));   // :208-209
	lastQuads = 0;
	if (frameCount > 0xffffffffull) throw std::runtime_error("more than 2^32 billboards in one frame");
	const uint32_t n = (uint32_t)frameCount;
	struct FrameReset { BillboardSystem &b; ~FrameReset() { b.pendingCount = 0; b.frameCount = 0; b.segments.clear(); } } frameReset { *this }; // :227-228
	if (!n) return;
	cudaStream_t st = ctx->stream;
	const uint32_t m = n < maxPerFrame ? n : maxPerFrame;
	const uint32_t numTiles = (n + 1023) / 1024;
	dBillboards.reserve(n, 0, st);
	for (int i = 0; i < 2; i++) { dKeys[i].reserve(n, 0, st); dVals[i].reserve(n, 0, st); }
	dTileHist.reserve((size_t)256 * numTiles, 0, st);
	dBlockCounts.reserve((m + 255) / 256, 0, st);
	dTotal.reserve(2, 0, st);
	dVerts.reserve((size_t)m * 24, 0, st);
	dAtlas.reserve(m, 0, st);
	// the frame's billboards in call order: converted runs out of the pinned staging array, packed arrays straight from
	// where the caller keeps them; a frame that is ONE packed device array is read in place
	const BillboardDev *frame = dBillboards.ptr;
	if (segments.size() == 1 && segments[0].kind == 2) frame = segments[0].src;
	else {
		size_t at = 0;
		for (const Segment &sg : segments) {
			const BillboardDev *src = sg.kind == 0 ? pending + sg.offset : sg.src;
			SPR_CUDA(cudaMemcpyAsync(dBillboards.ptr + at, src, sg.count * sizeof(BillboardDev), sg.kind == 2 ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
			at += sg.count;
		}
	}
	BillboardScratch sc;
	sc.keys[0] = dKeys[0].ptr; sc.keys[1] = dKeys[1].ptr; sc.vals[0] = dVals[0].ptr; sc.vals[1] = dVals[1].ptr;
	sc.tileHist = dTileHist.ptr; sc.blockCounts = dBlockCounts.ptr;
	dDraws.reserve(m, 0, st);
	ctx->launchCount += launch_billboards(st, frame, n, maxPerFrame, sc, dVerts.ptr, dAtlas.ptr, dTotal.ptr, dDraws.ptr);
	SPR_CUDA(cudaGetLastError());
	// the draw list (:198-203, runs of equal atlas) is built on the device: two counts, then one copy of 16 bytes per draw
	uint32_t totals[2] = { 0, 0 };
	SPR_CUDA(cudaMemcpyAsync(totals, dTotal.ptr, sizeof(totals), cudaMemcpyDeviceToHost, st));
	SPR_CUDA(cudaStreamSynchronize(st));
	const uint32_t total = totals[0], numDraws = totals[1];
	static_assert(sizeof(SprBillboardDraw) == sizeof(BillboardDrawDev), "SprBillboardDraw is BillboardDrawDev");
	drawsPinned.reserve(numDraws ? numDraws : 1);
	if (numDraws) {
		SPR_CUDA(cudaMemcpyAsync(drawsPinned.ptr, dDraws.ptr, (size_t)numDraws * sizeof(SprBillboardDraw), cudaMemcpyDeviceToHost, st));
		SPR_CUDA(cudaStreamSynchronize(st));
	}
	lastQuads = total;
	out.deviceVertices = (const SprBillboardVertex*)dVerts.ptr;
	out.numQuads = total;
	out.numDraws = numDraws;
	out.draws = drawsPinned.ptr;
}

// ---------------------------------------------------------------------------
// Area query
// ---------------------------------------------------------------------------

uint32_t ParticleSystem::queryEffects(const SprFrustum &frustum, uint32_t *outIds, uint32_t capacity)
{
	SPR_CUDA(cudaSetDevice(ctx->device));
	if (!ctx->dirtyParams.empty() || !ctx->initStateIdx.empty() || !ctx->typeIdx.empty() || !ctx->systemIdx.empty() || !ctx->ownerJobs.empty()) ctx->flushUploads();
	const uint32_t n = ctx->numEffectGlobals;
	if (!n) return 0;
	ctx->queryD.reserve((size_t)n + 1, 0, ctx->stream);
	SPR_CUDA(cudaMemsetAsync(ctx->queryD.ptr, 0, sizeof(uint32_t), ctx->stream));
	launch_query_frustum(ctx->stream, ctx->tablePtrs(), n, systemIdx, &frustum.side[0][0], ctx->queryD.ptr + 1, n, ctx->queryD.ptr);
	ctx->launchCount++;
	SPR_CUDA(cudaGetLastError());
	std::vector<uint32_t> host((size_t)n + 1);
	SPR_CUDA(cudaMemcpyAsync(host.data(), ctx->queryD.ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
	SPR_CUDA(cudaStreamSynchronize(ctx->stream));
	const uint32_t count = host[0];
	if (count) {
		SPR_CUDA(cudaMemcpyAsync(host.data() + 1, ctx->queryD.ptr + 1, (size_t)count * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
		SPR_CUDA(cudaStreamSynchronize(ctx->stream));
	}
	std::sort(host.begin() + 1, host.begin() + 1 + count); // ascending effect id: the defined order of this query
	for (uint32_t i = 0; i < count && i < capacity; i++) outIds[i] = host[1 + i];
	return count;
}

} // namespace spr
