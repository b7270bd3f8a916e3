// device_types.h -- structures shared by the host layer and the sm_100a kernels.
//
// Layout decisions (DESIGN.md "Data layout in HBM"):
//  * slot state is an AoS of 32-byte records {pos.xyz, life, vel.xyz, seed} (two float4),
//    i.e. one lane of the reference's Particle4 (ParticleSystem.cpp:170-176) per record;
//    slot index i is the only thing parity observes.
//  * all effects of all systems of a context live in ONE slot pool; an effect owns a
//    power-of-two range [slotBase, slotBase + slotCapacity) aligned to min(capacity, 1024).
//  * a 128-slot granule (one warp x 4 slots per lane) belongs to exactly one effect
//    (granuleOwner), a 1024-slot tile (one CTA) either belongs to one large effect or
//    holds whole small effects.
#pragma once

#include <stdint.h>

namespace spr {

static const uint32_t kGranuleSlots = 128;     // slots per warp
static const uint32_t kTileSlots = 1024;       // slots per CTA (8 warps)
static const uint32_t kMaxTimerSpawns = 20;    // ParticleSystem.cpp:478
static const uint32_t kNoOwner = 0xffffffffu;
#ifndef SPR_SORT_TILE_KEYS
#define SPR_SORT_TILE_KEYS 4096
#endif
static const uint32_t kSortTileKeys = SPR_SORT_TILE_KEYS; // keys per CTA in the radix sort passes (16 per thread)
// Forward progress of the chained scans (look-back over the status words of lower-numbered tiles).
//  * k_stream_fused is PERSISTENT (one wave of CTAs, several tiles each): its tiles are always handed out by an atomic
//    ticket, in the order the CTAs get to them, so a look-back only waits for tiles whose CTAs are running -- however
//    many of its CTAs the device can hold while other contexts / streams use it.
//  * k_sort_pass and k_dead_append launch ONE CTA (warp) PER TILE and take tile = block index. A tile waits only for
//    lower block indices of its own grid; the hardware work distributor hands the CTAs of a grid out in ascending linear
//    block index and never evicts a running CTA in favour of another one of the same grid (compute preemption swaps
//    whole contexts), so every tile waited for is running or finished, whatever else shares the device. That is the
//    dispatch order every non-ticketed decoupled look-back relies on; it is observed behaviour, not a documented
//    guarantee, so -DSPR_STRICT_TILE_TICKETS=1 builds these two kernels with tickets as well (same results; measured
//    on the 10M-particle sorted step: 0.777 -> 0.825 ms, the ticket's L2 round trip sits in front of every tile's
//    dependent loads -- profiles/r02_sort_experiments.md). A wait that outlasts 2^24 polls traps either way.
#ifndef SPR_STRICT_TILE_TICKETS
#define SPR_STRICT_TILE_TICKETS 0
#endif
static const uint32_t kSmallSortMax = 18432;   // an effect of at most this many slots is sorted by ONE CTA (k_sort_small)

// spline atlas of the vertex stage (ParticleSystem.cpp:253-258): 64 samples per curve, 8 x 256 type cells of 2 rows
static const uint32_t kSplineSamples = 64, kSplineAtlasTypesX = 8, kSplineAtlasTypesY = 256, kSplineRowsPerType = 2;
static const float kHugeLife = 1e20f;          // ParticleSystem.cpp:27
static const float kHugeLifeCmp = 1e19f;       // ParticleSystem.cpp:28

enum RandomVec3Flags { kRvBox = 1, kRvSphere = 2, kRvRotation = 4 }; // ParticleSystem.cpp:68-72

// RandomVec3Imp (ParticleSystem.cpp:66-116), baked on the host.
struct RandomVec3Dev {
	uint32_t flags;
	float offset[3];
	float boxExtent[3];
	float thetaBias, thetaScale;
	float cosPhiBias, cosPhiScale;
	float radiusCubeScale, radiusScale;
	float sphereScale[3];
	float rotation[4];
};

struct GravityPointDev {
	float position[3]; // emitter-local (ServerState.h:182)
	float radius;      // max(radius, 0.05), ParticleSystem.cpp:548
	float strength;    // negated, :549
};

// Per effect type constants the kernels read.
struct TypeDev {
	RandomVec3Dev emitPosition, emitVelocity;
	float attractorOffset[3];
	float attractorStrength;
	float spawnTime, spawnTimeVariance;
	uint32_t burstAmount;
	uint32_t drawsPerSpawn;           // K: RNG draws of one timer spawn; a burst spawn uses K-1
	uint64_t mulKm1, sumKm1;          // LCG jump by K-1 draws: s' = s*mulKm1 + inc*sumKm1
	uint64_t mulJK[kMaxTimerSpawns + 1], sumJK[kMaxTimerSpawns + 1]; // LCG jump by j*K draws, j = 0..20: the state in front of
	                                  // the j-th timer spawn of a sub-step (k_plan evaluates the <= 20 timer draws side by side)
	float drag, lifeTime, lifeTimeVariance /* pre-scaled by 2^-24, :535 */, cullPadding;
	uint32_t numGravityPoints;
	uint32_t localSpace;
	GravityPointDev gravityPoints[16];
};

// Host-owned per-effect parameters (uploaded when they change).
struct EffectParams {
	uint32_t typeIdx;       // context-wide type index
	uint32_t systemIdx;     // context-wide system index
	uint32_t slotBase;      // first pool slot
	uint32_t slotCapacity;  // power of two >= 128
	float timeStep;         // per-effect jittered step, ParticleSystem.cpp:744
	float gravity[3];
	float emitterToWorld[16]; // column-major 4x4, :200
	uint32_t hostStopEmit;  // prepareForRemove, :791
	uint32_t effectId;      // system-local id (what the area / visibility lists carry)
	uint32_t live;          // 0 once the effect is removed
	uint32_t _pad;
	float areaSphere[4];    // the effect's sphere area: transform position + updateRadius (:762-763, :782-783)
};

// Device-owned per-effect state.
struct EffectState {
	float prevEmitterToWorld[16]; // :201
	float spawnTimer;             // :210
	float emitOnTimer;            // :211
	uint32_t stopEmit;            // :207 (device part: emitOnTimer expiry)
	uint32_t firstEmit;           // :208
	uint32_t parity;              // which entry of slotCount/freeCount is current
	uint32_t slotCount[2];        // particles.size * 4
	uint32_t freeCount[2];        // freeIndices.size
	uint32_t aliveCount;          // gpuParticles.size / 4
	uint32_t boundsMin[4];        // order-preserving uint encoding of min(px,py,pz,life), :532
	uint32_t boundsMax[4];
	uint32_t keyMin[8];           // sort mode: smallest depth key of the step = min over the 8 words (one per warp position
	                              // in the CTA, so that the first wave of atomics does not pile up on one address);
	                              // the radix sort runs on key - keyMin
};

__host__ __device__ inline uint32_t effect_key_min(const EffectState &S)
{
	uint32_t m = S.keyMin[0];
	for (int i = 1; i < 8; i++) m = S.keyMin[i] < m ? S.keyMin[i] : m;
	return m;
}

#ifdef __CUDACC__
// Depth key of a live particle: ~enc(|camera - p|^2), sf::lengthSq (src/sf/Vector.h:99) in its operation order; an
// ascending stable sort of the keys is back to front with ties by slot (BillboardSystem.cpp:41-51). ONE definition for
// every kernel that needs the key of a record.
__device__ __forceinline__ uint32_t depth_key(float px, float py, float pz, float camX, float camY, float camZ)
{
	const float dx = camX - px, dy = camY - py, dz = camZ - pz;
	const uint32_t u = __float_as_uint(dx*dx + dy*dy + dz*dz);
	return ~((u & 0x80000000u) ? ~u : (u | 0x80000000u));
}
#endif

struct SystemDev {
	uint64_t rngState, rngInc;    // updateCtx.rng, :251,:267
};

// One system's share of a frame.
struct SystemWork {
	uint32_t systemIdx;
	uint32_t firstActive;   // index into the frame's active arrays
	uint32_t numActive;
	uint32_t _pad;
	float camera[3];        // sort mode: the request's FrameArgs::mainRenderArgs.cameraPosition (depth keys of this system's particles)
	uint32_t _pad2;
};

// One active effect of the frame (host computed: sub-step count from timeDelta, :816-820).
struct ActiveEntry {
	uint32_t effect;        // context-wide effect index
	uint32_t numSubsteps;
	uint32_t planBase;      // first PlanRec of this effect in the frame
	uint32_t flags;         // bit0: burst handled by the wide burst kernel; bit1: the host wants the exact slot count mirrored (its capacity bound is not final);
	                        // [31:2]: index of the effect's system in the frame's SystemWork list
};

// Spawn plan of one simulateParticlesImp call (one effect, one sub-step), written by
// the planner kernel (control flow of ParticleSystem.cpp:457-485,:524).
struct PlanRec {
	uint64_t rngState;      // RNG state before the first draw of this sub-step
	uint32_t effect;
	uint32_t burst;         // burst spawns (all precede timer spawns, :480)
	uint32_t timer;         // timer spawns (<= 20)
	uint32_t flags;         // bit0 usePrev (prevEmitterToWorld is the stored one), bit1 count parity in
	float dt;
	float spawnTimer0;      // spawnTimer after `-= dt` (:477): emitT of burst spawns
	float timerVals[kMaxTimerSpawns]; // spawnTimer after each timer increment (:484)
};

// Everything a slot needs to know about its effect in one sub-step, digested by k_emit_warp into
// one 80-byte record so that the streaming kernels reach their data loads after two dependent
// loads (granule owner -> this record) instead of seven.
struct __align__(16) RoundConsts {
	uint32_t stamp, round;   // valid for this (frame stamp, sub-step round) only
	uint32_t slotBase;
	uint32_t slotsPost;      // particles.size*4 after the emission of this sub-step
	uint32_t topPost;        // freeIndices.size after the emission of this sub-step
	uint32_t flags;          // bit0 parity in, bit1 large (capacity > 1024), bit2 usePrev, bit3 grew (nSpawn > topIn), [15:8] numGravityPoints
	uint32_t typeIdx, planIdx;
	float dt, drag, lifeTime, lifeVar;
	float gx, gy, gz;
	uint32_t _pad;
	float camX, camY, camZ; // sort mode: the camera of the effect's system in this frame (depth keys)
	uint32_t _pad2;
};
static_assert(sizeof(RoundConsts) == 80, "RoundConsts is read as five 16-byte words");

// A job of the wide burst-emission kernel: `count` burst spawns of active entry `entry`
// starting at CTA `firstBlock`.
struct BurstJob {
	uint32_t entry;
	uint32_t firstBlock;
	uint32_t numBlocks;
	uint32_t _pad;
};

struct FillJob { uint32_t start, count, value, _pad; };

// One draw record for the vertex-stage kernel (vertex_stage.cu): `numParticles` quads read from stream record
// `srcRecord` (32-byte units), written from vertex `dstVertex` on, by CTAs [firstBlock, firstBlock + ceil(n / 256)).
struct ShadeJob {
	uint64_t srcRecord;
	uint64_t dstVertex;
	uint32_t numParticles;
	uint32_t lutIndex;        // which 128-word LUT of the call's LUT table (one per distinct effect type)
	uint32_t firstBlock;
	uint32_t _pad;
	float instance[40];       // Particle_VertexInstance_t, 160 bytes (ParticleSystem.cpp:871-876)
	float type[24];           // Particle_VertexType_t, 96 bytes (:312-321, :410-419)
};

// Sort tile: `tile`-th 4096-key tile of effect `effect`'s pair segment.
struct SortTile { uint32_t effect; uint32_t tile; };

// One billboard of the frame (cl::BillboardSystemImp::BillboardImp + its BillboardOrder::depth,
// src/client/BillboardSystem.cpp:30-51, with the sp::Sprite fields renderMain reads, src/sp/Sprite.h:44-57).
struct BillboardDev {
	unsigned long long atlas;   // sprite->atlas (opaque id)
	float minVert[2], maxVert[2], vertUvScale[2], vertUvBias[2];
	float transform[12];        // sf::Mat34, column major (src/sf/Matrix.h:86-97)
	float anchor[2];
	uint32_t color;             // packColor(), :17-28 (packed on the host at add time, like the reference)
	float cropMin[2], cropMax[2];
	float depth;
};

// One sg_draw of cl::BillboardSystem::renderMain (:198-203): `count` quads of one atlas from quad `firstQuad` on (= SprBillboardDraw)
struct BillboardDrawDev { unsigned long long atlas; uint32_t count, firstQuad; };

// Launch-constant pointers for the per-round kernels.
struct PoolPtrs {
	float4 *state;            // 2 float4 per slot
	uint32_t *freeStack;      // per slot capacity; effect e uses [slotBase, slotBase+freeCount)
	float4 *vertices;         // 8 float4 per slot capacity (4 x 32 B per live particle)
	uint32_t *keys;           // sort mode: depth keys (descending-order transform applied); written per SLOT by
	                          // k_integrate_sort, compacted by the first radix pass
	uint32_t *vals;           // sort mode: local slot indices, compacted/sorted alongside the keys
	uint32_t *aliveMask;      // one word per 32 slots, bit = slot emits a particle this step (:611)
	uint32_t *deadMask;       // one word per 32 slots, bit = slot was freed this step (:556-567)
	uint32_t *wordAliveBase;  // stream mode: rank in the effect's stream of the first live slot of each 32-slot word
	const uint32_t *granuleOwner;
	unsigned long long *tileStatusAlive;
	unsigned long long *tileStatusDead;
	unsigned long long *tickets;      // [2] monotonic tile tickets of the chained-scan kernels (0: k_stream_fused, 1: k_dead_append):
	                                  // a CTA / warp takes its tile in the order it STARTS, so a look-back only ever waits for
	                                  // tiles whose owners are already running, whatever else occupies the device
};

struct TablePtrs {
	const EffectParams *params;
	EffectState *state;
	const TypeDev *types;
	SystemDev *systems;
	PlanRec *plans;
	RoundConsts *roundConsts; // per effect, rewritten every sub-step round by k_emit_warp
	uint32_t *effectAnyDead;  // per effect: some slot was freed in this round (lets k_dead_append skip its scan)
	unsigned long long *slotMirror; // per effect, in mapped host memory: frame stamp << 32 | slot count after the frame
};

} // namespace spr
