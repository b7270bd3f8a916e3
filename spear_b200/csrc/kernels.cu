// kernels.cu -- hand-written sm_100a kernels for the particle hot path.
//
// Reference semantics: src/client/ParticleSystem.cpp:452-680 (simulateParticlesImp).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (bit-exact float
// parity needs every multiply-add to round twice).
//
// Per frame (one set of launches covers every system of the context):
//   k_plan              one warp per system: spawn control flow + RNG stream offsets
//   per sub-step round:
//     k_emit_warp       one warp per active effect: timer spawns (+ small bursts), RoundConsts digest
//     k_emit_burst      one thread per burst spawn of large bursts
//     k_integrate_pass  one warp per 128-slot granule, pure streaming: integrate, ballots, (depth key),
//                       bounds; single-granule effects are finished in the warp
//     k_dead_append     one warp per 1024-slot tile over the ballots: free-list append, counts,
//                       (stream mode) stream ranks; chained look-back across the tiles of a large effect
//   stream mode:        k_expand_stream   4x vertex expansion in slot order
//   sort mode:          sort.cu           segmented radix sort + k_expand_sorted
//   k_finalize          prevEmitterToWorld = emitterToWorld (:668)
//
// Everything order-sensitive (compaction order, free-list order, spawn->slot
// assignment, RNG offsets) is a scan, never an atomic.

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "device_types.h"
#include "kernels.h"
#include "launch.h"

namespace spr {

#define FULL_MASK 0xffffffffu

// ---------------------------------------------------------------------------
// PCG-XSH-RR 64/32 (src/sf/Random.cpp:8-20) with LCG jump-ahead
// ---------------------------------------------------------------------------

static const uint64_t kPcgMul = 6364136223846793005ULL;

__host__ __device__ inline uint32_t pcg_output(uint64_t old)
{
	uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
	uint32_t rot = (uint32_t)(old >> 59u);
	return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31u));
}

struct Pcg {
	uint64_t state, inc;
	__device__ inline uint32_t nextU32() { uint64_t old = state; state = old * kPcgMul + inc; return pcg_output(old); }
	// (float)((double)u * 2^-32) == RN(u) * 2^-32 exactly
	__device__ inline float nextFloat() { return __uint2float_rn(nextU32()) * 2.3283064365386963e-10f; }
	// advance by `delta` draws in O(log delta)
	__device__ inline void advance(uint64_t delta)
	{
		uint64_t curMul = kPcgMul, curPlus = inc, accMul = 1, accPlus = 0;
		while (delta > 0) {
			if (delta & 1) { accMul *= curMul; accPlus = accPlus * curMul + curPlus; }
			curPlus = (curMul + 1) * curPlus;
			curMul *= curMul;
			delta >>= 1;
		}
		state = accMul * state + accPlus;
	}
};

// ---------------------------------------------------------------------------
// Approximations and samplers (ParticleSystem.cpp:43-64, :118-145)
// ---------------------------------------------------------------------------

__device__ inline float sf_min(float a, float b) { return a < b ? a : b; } // minss
__device__ inline float sf_max(float a, float b) { return a > b ? a : b; } // maxss

__device__ inline float approx_acos(float a)
{
	float x = fabsf(a);
	float v = -1.280827681872808f + 1.280827681872808f * __fsqrt_rn(1.0f - x) - 0.28996864492208857f * x;
	return 3.14159265358979323846f * 0.5f - copysignf(v, a);
}

__device__ inline float cossin_lane(float v, float bias)
{
	const float rcp2pi = 1.0f / 6.28318530717958647692f;
	float t = v;
	t *= rcp2pi;
	t += bias;
	t -= rintf(t - 0.25f) + 0.25f;
	t *= (fabsf(t) - 0.5f) * 16.0f;
	t += t * (fabsf(t) - 1.0f) * 0.225f;
	return t;
}

__device__ inline float approx_cbrt(float a)
{
	return 1.5142669123810535f * __fsqrt_rn(a) - 0.5142669123810535f * a;
}

__device__ inline float3 sample_random_vec3(const RandomVec3Dev &r, Pcg &rng)
{
	float3 p = make_float3(r.offset[0], r.offset[1], r.offset[2]);
	if (r.flags & kRvBox) {
		float x = rng.nextFloat(), y = rng.nextFloat(), z = rng.nextFloat();
		p.x += (x - 0.5f) * r.boxExtent[0];
		p.y += (y - 0.5f) * r.boxExtent[1];
		p.z += (z - 0.5f) * r.boxExtent[2];
	}
	if (r.flags & kRvSphere) {
		float u = rng.nextFloat();
		float v = rng.nextFloat();
		float w = rng.nextFloat();
		float theta = r.thetaBias + u * r.thetaScale;
		float phi = approx_acos(r.cosPhiBias + v * r.cosPhiScale);
		float cosTheta = cossin_lane(theta, 0.0f), sinTheta = cossin_lane(theta, -0.25f);
		float cosPhi = cossin_lane(phi, 0.0f), sinPhi = cossin_lane(phi, -0.25f);
		float radius = approx_cbrt(1.0f - w * r.radiusCubeScale) * r.radiusScale;
		p.x += (sinPhi * cosTheta) * radius * r.sphereScale[0];
		p.y += (cosPhi) * radius * r.sphereScale[1];
		p.z += (sinPhi * sinTheta) * radius * r.sphereScale[2];
	}
	if (r.flags & kRvRotation) { // sf::rotate, src/sf/Quaternion.h:66-75
		float qx = r.rotation[0], qy = r.rotation[1], qz = r.rotation[2], qw = r.rotation[3];
		float xy = qx*p.y - qy*p.x;
		float xz = qx*p.z - qz*p.x;
		float yz = qy*p.z - qz*p.y;
		float3 o;
		o.x = 2.0f * (+ qw*yz + qy*xy + qz*xz) + p.x;
		o.y = 2.0f * (- qx*xy - qw*xz + qz*yz) + p.y;
		o.z = 2.0f * (- qx*xz - qy*yz + qw*xy) + p.z;
		p = o;
	}
	return p;
}

// order-preserving float <-> uint (min/max through integer atomics and redux)
__device__ inline uint32_t enc_float(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }

// ---------------------------------------------------------------------------
// Small utility kernels
// ---------------------------------------------------------------------------

__global__ void k_fill_jobs(uint32_t *dst, const FillJob *jobs, uint32_t numJobs)
{
	pdl_enter();
	uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	uint32_t numWarps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t j = warp; j < numJobs; j += numWarps) {
		FillJob job = jobs[j];
		for (uint32_t i = lane; i < job.count; i += 32) dst[job.start + i] = job.value;
	}
}

// dst[idx[i]] = src[i] for structs of `words` 32-bit words
__global__ void k_scatter_words(uint32_t *dst, const uint32_t *idx, const uint32_t *src, uint32_t n, uint32_t words)
{
	pdl_enter();
	uint64_t total = (uint64_t)n * words;
	for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t i = (uint32_t)(t / words), w = (uint32_t)(t % words);
		dst[(uint64_t)idx[i] * words + w] = src[t];
	}
}

// ---------------------------------------------------------------------------
// k_plan: spawn control flow of simulateParticlesImp (:457-485, :524) for every
// (effect, sub-step) of every system, in the reference's RNG consumption order
// (all sub-steps of effect A before effect B, :812-821).
// ---------------------------------------------------------------------------

// One WARP per system. The RNG chain makes the effects of a system sequential, but only the few
// arithmetic steps of the spawn loop are: each lane first loads everything one effect needs (32
// effects' dependent global loads in flight at once), then the effects that draw take turns in list order -- the whole
// warp runs a turn, the timer draws of a sub-step side by side, the RNG state it leaves is the next turn's -- and
// finally every lane writes its own effect's state back in parallel.
__global__ void __launch_bounds__(128) k_plan(TablePtrs tab, const SystemWork *work, uint32_t numWork, const ActiveEntry *active, uint32_t stamp)
{
	pdl_enter();
	const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (w >= numWork) return;
	const SystemWork sw = work[w];
	Pcg rng;
	rng.state = tab.systems[sw.systemIdx].rngState;
	rng.inc = tab.systems[sw.systemIdx].rngInc;

	for (uint32_t base = 0; base < sw.numActive; base += 32) {
		const uint32_t a = base + lane;
		const bool have = a < sw.numActive;
		ActiveEntry ae = { 0, 0, 0, 0 };
		float spawnTimer = 0.0f, emitOnTimer = 0.0f, dt = 0.0f, spawnTime = 0.0f, spawnVar = 0.0f;
		uint32_t stopEmit = 0, firstEmit = 0, parity = 0, hostStop = 0, K = 2, burstAmount = 0;
		uint64_t mulKm1 = 1, sumKm1 = 0;
		if (have) {
			ae = active[sw.firstActive + a];
			const EffectParams &P = tab.params[ae.effect];
			const EffectState &S = tab.state[ae.effect];
			const TypeDev &T = tab.types[P.typeIdx];
			spawnTimer = S.spawnTimer; emitOnTimer = S.emitOnTimer;
			stopEmit = S.stopEmit; firstEmit = S.firstEmit; parity = S.parity;
			dt = P.timeStep; hostStop = P.hostStopEmit;
			K = T.drawsPerSpawn; burstAmount = T.burstAmount;
			spawnTime = T.spawnTime; spawnVar = T.spawnTimeVariance;
			mulKm1 = T.mulKm1; sumKm1 = T.sumKm1;
		}
		// The control loop of one effect (:457-485, :524). `rng` is only touched when something spawns.
		auto run_effect = [&]() {
			for (uint32_t s = 0; s < ae.numSubsteps; s++) {
				PlanRec &rec = tab.plans[ae.planBase + s];
				uint32_t burst = 0, usePrev = (s == 0) ? 1u : 0u;
				if (firstEmit) { burst = burstAmount; firstEmit = 0; usePrev = 0; }     // :457-464
				if (emitOnTimer >= 0.0f) {                                              // :466-471
					emitOnTimer -= dt;
					if (emitOnTimer < 0.0f) stopEmit = 1;
				}
				const bool stop = stopEmit || hostStop;
				spawnTimer -= dt;                                                        // :477
				rec.rngState = rng.state;
				rec.effect = ae.effect;
				rec.burst = burst;
				rec.dt = dt;
				rec.spawnTimer0 = spawnTimer;
				rec.flags = usePrev | ((parity & 1u) << 1);
				if (burst) rng.advance((uint64_t)burst * (K - 1));                       // burst spawns draw first (:480-481)
				uint32_t timer = 0;
				int spawnsLeft = (int)kMaxTimerSpawns;
				while (spawnTimer <= 0.0f && !stop) {                                    // :479-485
					if (spawnsLeft-- <= 0) break;
					spawnTimer += spawnTime + spawnVar * rng.nextFloat();
					rec.timerVals[timer++] = spawnTimer;
					rng.state = rng.state * mulKm1 + rng.inc * sumKm1;                   // the spawn's own K-1 draws
				}
				rec.timer = timer;
				spawnTimer = sf_max(spawnTimer, 0.0f);                                   // :524
				parity ^= 1u;
			}
		};
		// Most effects of a frame spawn nothing (the burst is over, the timer has not expired) and then draw nothing:
		// a dry run of the same conditions on copies of the state tells, and those effects run their loop right away,
		// all lanes at once (the rngState they record is never read: no spawn uses it). Only the effects that do
		// spawn take turns on the RNG chain, in list order.
		bool draws = false;
		if (have) {
			float st = spawnTimer, eo = emitOnTimer;
			uint32_t se = stopEmit, fe = firstEmit;
			for (uint32_t s = 0; s < ae.numSubsteps && !draws; s++) {
				if (fe) { fe = 0; if (burstAmount) draws = true; }
				if (eo >= 0.0f) { eo -= dt; if (eo < 0.0f) se = 1; }
				st -= dt;
				if (st <= 0.0f && !(se || hostStop)) draws = true;
				st = sf_max(st, 0.0f);
			}
			if (!draws) run_effect();
		}
		// The effects that draw take turns, in list order, but the whole WARP runs each turn: all lanes hold the turn's
		// effect (broadcast), lane j evaluates the j-th timer draw of the sub-step from the state jumped ahead by j * K
		// draws (TypeDev::mulJK), and only the <= 20 float additions that decide how many of them happen stay sequential --
		// instead of one lane walking 20 dependent 64-bit multiply-adds per sub-step while 31 wait (C5: 100 effects x 20
		// spawns per system and frame).
		uint32_t turn = __ballot_sync(FULL_MASK, draws);
		const uint32_t myTypeIdx = have ? tab.params[ae.effect].typeIdx : 0u;
		while (turn) {
			const uint32_t k = __ffs(turn) - 1;
			turn &= turn - 1;
			// the turn's effect, in every lane
			const uint32_t tEffect = __shfl_sync(FULL_MASK, ae.effect, k), tSub = __shfl_sync(FULL_MASK, ae.numSubsteps, k), tPlan = __shfl_sync(FULL_MASK, ae.planBase, k);
			float tSpawnTimer = __shfl_sync(FULL_MASK, spawnTimer, k), tEmitOn = __shfl_sync(FULL_MASK, emitOnTimer, k);
			uint32_t tStopEmit = __shfl_sync(FULL_MASK, stopEmit, k), tFirst = __shfl_sync(FULL_MASK, firstEmit, k), tParity = __shfl_sync(FULL_MASK, parity, k);
			const uint32_t tHostStop = __shfl_sync(FULL_MASK, hostStop, k), tK = __shfl_sync(FULL_MASK, K, k), tBurstAmount = __shfl_sync(FULL_MASK, burstAmount, k);
			const float tDt = __shfl_sync(FULL_MASK, dt, k), tSpawnTime = __shfl_sync(FULL_MASK, spawnTime, k), tSpawnVar = __shfl_sync(FULL_MASK, spawnVar, k);
			const TypeDev &TT = tab.types[__shfl_sync(FULL_MASK, myTypeIdx, k)];
			const uint32_t jl = lane <= kMaxTimerSpawns ? lane : kMaxTimerSpawns;
			const uint64_t mulJ = TT.mulJK[jl], sumJ = TT.sumJK[jl];
			for (uint32_t s = 0; s < tSub; s++) {
				PlanRec &rec = tab.plans[tPlan + s];
				uint32_t burst = 0, usePrev = (s == 0) ? 1u : 0u;
				if (tFirst) { burst = tBurstAmount; tFirst = 0; usePrev = 0; }          // :457-464
				if (tEmitOn >= 0.0f) {                                                  // :466-471
					tEmitOn -= tDt;
					if (tEmitOn < 0.0f) tStopEmit = 1;
				}
				const bool stop = tStopEmit || tHostStop;
				tSpawnTimer -= tDt;                                                      // :477
				if (lane == 0) {
					rec.rngState = rng.state;
					rec.effect = tEffect;
					rec.burst = burst;
					rec.dt = tDt;
					rec.spawnTimer0 = tSpawnTimer;
					rec.flags = usePrev | ((tParity & 1u) << 1);
				}
				if (burst) rng.advance((uint64_t)burst * (tK - 1));                      // burst spawns draw first (:480-481)
				// lane j: the draw of the j-th timer spawn, were it to happen (:484): state jumped by j * K draws
				const float myInc = tSpawnTime + tSpawnVar * (__uint2float_rn(pcg_output(rng.state * mulJ + rng.inc * sumJ)) * 2.3283064365386963e-10f);
				uint32_t timer = 0;
				if (tSpawnTimer <= 0.0f && !stop) {
					float inc[kMaxTimerSpawns];
#pragma unroll
					for (uint32_t j = 0; j < kMaxTimerSpawns; j++) inc[j] = __shfl_sync(FULL_MASK, myInc, j);
					// :479-485, at most 20 spawns: spawn j happens while the timer is <= 0. The chain of additions runs
					// unconditionally (the values behind the last spawn are simply not used), so that only the 20 float
					// additions depend on one another, not 20 x (compare -> add).
					float mine = 0.0f, t = tSpawnTimer;
					bool going = true;
#pragma unroll
					for (uint32_t j = 0; j < kMaxTimerSpawns; j++) {
						going = going && t <= 0.0f;
						t += inc[j];
						if (going) { tSpawnTimer = t; timer = j + 1; }
						if (lane == j) mine = t;
					}
					if (lane < timer) rec.timerVals[lane] = mine;
					// the state behind `timer` spawns of K draws each
					const uint64_t mulN = __shfl_sync(FULL_MASK, mulJ, timer), sumN = __shfl_sync(FULL_MASK, sumJ, timer);
					rng.state = rng.state * mulN + rng.inc * sumN;
				}
				if (lane == 0) rec.timer = timer;
				tSpawnTimer = sf_max(tSpawnTimer, 0.0f);                                 // :524
				tParity ^= 1u;
			}
			if (lane == k) { spawnTimer = tSpawnTimer; emitOnTimer = tEmitOn; stopEmit = tStopEmit; firstEmit = tFirst; parity = tParity; }
		}
		if (have && ae.numSubsteps) {
			EffectState &S = tab.state[ae.effect];
			S.spawnTimer = spawnTimer;
			S.emitOnTimer = emitOnTimer;
			S.stopEmit = stopEmit;
			S.firstEmit = firstEmit;
			S.parity = parity;
		}
	}
	if (lane == 0) tab.systems[sw.systemIdx].rngState = rng.state;
}

// ---------------------------------------------------------------------------
// Emission (:487-522)
// ---------------------------------------------------------------------------

struct EmitCtx {
	const EffectParams *P;
	const EffectState *S;
	const TypeDev *T;
	const PlanRec *rec;
	uint64_t inc;
	uint32_t top, blocks; // free-list size and Particle4 block count before this sub-step
	float e[16], pe[16];  // emitterToWorld / prevEmitterToWorld
};

__device__ inline void emit_ctx_init(EmitCtx &c, const TablePtrs &tab, const PlanRec *rec)
{
	c.rec = rec;
	uint32_t g = rec->effect;
	c.P = &tab.params[g];
	c.S = &tab.state[g];
	c.T = &tab.types[c.P->typeIdx];
	c.inc = tab.systems[c.P->systemIdx].rngInc;
	uint32_t par = (rec->flags >> 1) & 1u;
	c.top = c.S->freeCount[par];
	c.blocks = c.S->slotCount[par] >> 2;
#pragma unroll
	for (int i = 0; i < 16; i++) {
		c.e[i] = c.P->emitterToWorld[i];
		c.pe[i] = (rec->flags & 1u) ? c.S->prevEmitterToWorld[i] : c.e[i];
	}
}

// Spawn number j of the sub-step (burst spawns first, then timer spawns).
__device__ inline void emit_spawn(const EmitCtx &c, const PoolPtrs &pool, uint32_t j)
{
	const PlanRec &rec = *c.rec;
	const TypeDev &T = *c.T;
	const uint32_t K = T.drawsPerSpawn;
	Pcg rng;
	rng.state = rec.rngState;
	rng.inc = c.inc;
	float timerVal;
	if (j < rec.burst) {
		rng.advance((uint64_t)j * (K - 1));
		timerVal = rec.spawnTimer0;
	} else {
		uint32_t k = j - rec.burst;
		rng.advance((uint64_t)rec.burst * (K - 1) + (uint64_t)k * K + 1); // +1: the timer draw (:484)
		timerVal = rec.timerVals[k];
	}

	// slot: pop of the LIFO free list, growing by one Particle4 at a time (:487-496)
	uint32_t slot;
	if (j < c.top) {
		slot = pool.freeStack[c.P->slotBase + (c.top - 1 - j)];
	} else {
		uint32_t r = j - c.top;
		slot = 4u * (c.blocks + (r >> 2)) + 3u - (r & 3u);
	}

	float3 pos = sample_random_vec3(T.emitPosition, rng);   // :498
	float3 vel = sample_random_vec3(T.emitVelocity, rng);   // :499

	if (T.attractorStrength != 0.0f) {                      // :501-503
		float dx = pos.x - T.attractorOffset[0], dy = pos.y - T.attractorOffset[1], dz = pos.z - T.attractorOffset[2];
		float len = __fsqrt_rn(dx*dx + dy*dy + dz*dz);
		if ((double)len > 1e-9) {
			float rcp = __fdiv_rn(1.0f, len);
			dx *= rcp; dy *= rcp; dz *= rcp;
		} else {
			dx = dy = dz = 0.0f;
		}
		vel.x += dx * T.attractorStrength;
		vel.y += dy * T.attractorStrength;
		vel.z += dz * T.attractorStrength;
	}

	float emitT = sf_min(sf_max(__fdiv_rn(-timerVal, rec.dt), 0.0f), 1.0f); // :505
	float world[3];
#pragma unroll
	for (int k = 0; k < 3; k++) {                           // :506-509
		float p = c.e[12 + k] + (c.pe[12 + k] - c.e[12 + k]) * emitT;
		p += (c.e[0 + k] + (c.pe[0 + k] - c.e[0 + k]) * emitT) * pos.x;
		p += (c.e[4 + k] + (c.pe[4 + k] - c.e[4 + k]) * emitT) * pos.y;
		p += (c.e[8 + k] + (c.pe[8 + k] - c.e[8 + k]) * emitT) * pos.z;
		world[k] = p;
	}
	float seed = rng.nextFloat() * 16777216.0f;             // :522

	float4 *dst = pool.state + 2ull * (c.P->slotBase + slot);
	dst[0] = make_float4(world[0], world[1], world[2], 1.0f);
	dst[1] = make_float4(vel.x, vel.y, vel.z, seed);
}

// The last spawn of a sub-step that grew the effect initialises the unused lanes of the
// final new Particle4 (life = 1e20, :490); their free-list entries are written by k_integrate.
__device__ inline void emit_init_leftover(const EmitCtx &c, const PoolPtrs &pool, uint32_t n)
{
	if (n <= c.top) return;
	uint32_t nNew = n - c.top;
	uint32_t newBlocks = (nNew + 3) >> 2;
	uint32_t leftover = newBlocks * 4 - nNew;
	uint32_t lastBlock = c.blocks + newBlocks - 1;
	for (uint32_t i = 0; i < leftover; i++) {
		float4 *dst = pool.state + 2ull * (c.P->slotBase + 4 * lastBlock + i);
		dst[0] = make_float4(0.0f, 0.0f, 0.0f, kHugeLife);
		dst[1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	}
}

// One warp per active effect: all timer spawns, and the burst unless it is flagged for
// the wide kernel. Also resets the per-step bounds and handles the empty-effect corner.
__global__ void __launch_bounds__(256) k_emit_warp(PoolPtrs pool, TablePtrs tab, const SystemWork *work, const ActiveEntry *active, uint32_t numActive, uint32_t round, uint32_t stamp)
{
	pdl_launch_dependents();
	uint32_t a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (a >= numActive) return;
	ActiveEntry ae = active[a];               // the entry list came with the frame's host-to-device copy: read (and left) ahead of the wait
	if (round >= ae.numSubsteps) return;
	pdl_wait();
	const PlanRec *rec = &tab.plans[ae.planBase + round];
	EmitCtx c;
	emit_ctx_init(c, tab, rec);
	uint32_t n = rec->burst + rec->timer;
	uint32_t first = (ae.flags & 1u) ? rec->burst : 0u;
	for (uint32_t j = first + lane; j < n; j += 32) emit_spawn(c, pool, j);
	if (lane == 0) {
		EffectState &S = tab.state[rec->effect];
		emit_init_leftover(c, pool, n);
		// pMin = +HUGE_VALF, pMax = -HUGE_VALF (:532)
		uint32_t pinf = 0xff800000u /* enc(+inf) */, ninf = 0x007fffffu /* enc(-inf) */;
#pragma unroll
		for (int k = 0; k < 4; k++) { S.boundsMin[k] = pinf; S.boundsMax[k] = ninf; }
		for (int k = 0; k < 8; k++) S.keyMin[k] = 0xffffffffu;
		uint32_t slots = c.blocks * 4, topPost = c.top - (n <= c.top ? n : c.top);
		if (n > c.top) { uint32_t nb = (n - c.top + 3) >> 2; slots += 4 * nb; topPost = nb * 4 - (n - c.top); }
		uint32_t par = (rec->flags >> 1) & 1u;
		// the host sizes every range from a conservative (or, for short-lived effects, proven) slot bound: an effect
		// that outgrows it would write into its neighbour's slots -- stop the context instead
		if (slots > c.P->slotCapacity) __trap();
		if (slots == 0) { // nothing for k_integrate to do: carry the counts over
			S.slotCount[par ^ 1u] = 0;
			S.freeCount[par ^ 1u] = 0;
			S.aliveCount = 0;
		}
		RoundConsts rc;
		rc.stamp = stamp; rc.round = round;
		rc.slotBase = c.P->slotBase;
		rc.slotsPost = slots;
		rc.topPost = topPost;
		rc.flags = par | (c.P->slotCapacity > kTileSlots ? 2u : 0u) | ((rec->flags & 1u) << 2) | (n > c.top ? 8u : 0u) | (c.P->slotCapacity == kGranuleSlots ? 16u : 0u) | (c.T->numGravityPoints << 8);
		rc.typeIdx = c.P->typeIdx; rc.planIdx = ae.planBase + round;
		rc.dt = rec->dt; rc.drag = c.T->drag; rc.lifeTime = c.T->lifeTime; rc.lifeVar = c.T->lifeTimeVariance;
		rc.gx = c.P->gravity[0]; rc.gy = c.P->gravity[1]; rc.gz = c.P->gravity[2];
		rc._pad = 0;
		const SystemWork &sw = work[ae.flags >> 2];
		rc.camX = sw.camera[0]; rc.camY = sw.camera[1]; rc.camZ = sw.camera[2];
		rc._pad2 = 0;
		tab.roundConsts[rec->effect] = rc;
		tab.effectAnyDead[rec->effect] = 0;
	}
}

// One thread per burst spawn, for bursts too large for one warp.
__global__ void __launch_bounds__(256) k_emit_burst(PoolPtrs pool, TablePtrs tab, const ActiveEntry *active, const BurstJob *jobs, uint32_t numJobs)
{
	pdl_enter();
	// find the job of this CTA (jobs are sorted by firstBlock; numJobs is small)
	uint32_t lo = 0, hi = numJobs;
	while (hi - lo > 1) {
		uint32_t mid = (lo + hi) >> 1;
		if (jobs[mid].firstBlock <= blockIdx.x) lo = mid; else hi = mid;
	}
	BurstJob job = jobs[lo];
	ActiveEntry ae = active[job.entry];
	const PlanRec *rec = &tab.plans[ae.planBase];
	uint32_t j = (blockIdx.x - job.firstBlock) * blockDim.x + threadIdx.x;
	if (j >= rec->burst) return;
	EmitCtx c;
	emit_ctx_init(c, tab, rec);
	emit_spawn(c, pool, j);
}

// ---------------------------------------------------------------------------
// k_integrate (:526-666)
// ---------------------------------------------------------------------------

// A chained scan only ever waits for a predecessor that is already running: tiles are handed out by an atomic ticket in
// the order their CTAs (warps) start, never by block index, so the owner of every lower-numbered tile is resident
// whatever else shares the device (other contexts, other streams, a programmatically launched successor). A wait is a
// few microseconds; the trap after ~2^24 polls only turns a bug into an error the host reports instead of a hang.
static const uint32_t kMaxSpins = 1u << 24;
__device__ __forceinline__ void spin_guard(uint32_t &spins) { if (++spins > kMaxSpins) __trap(); }

__device__ inline unsigned long long ld_status(const unsigned long long *p)
{
	unsigned long long v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ inline void st_status(unsigned long long *p, unsigned long long v)
{
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// One slot record {pos.xyz, life, vel.xyz, seed} = 32 bytes = one DRAM sector: moved with a single
// 256-bit access per lane (LDG.E.256 / STG.E.256 on sm_100a), so a warp request covers 1 KB
// contiguous and every sector is touched exactly once.
__device__ inline void ld_record(const float4 *p, float4 &a, float4 &b)
{
	asm("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ inline void st_record(float4 *p, const float4 &a, const float4 &b)
{
	asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
		:: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// Per-effect constants of one sub-step, as every slot of the effect needs them.
struct StepConsts {
	const PlanRec *rec; const EffectParams *P; const TypeDev *T; EffectState *S;
	float dt, drag, lifeTime, lifeVar, gx, gy, gz;
	uint32_t numGp;
};

// Gravity points (:539-550 world transform, :581-590 force). Rare in practice: the kernels are
// instantiated with and without this path (GP) and the host picks the lean one when no live
// effect type has gravity points, which saves the streaming kernels ~20 registers.
__device__ __forceinline__ void gravity_points_accel(const PlanRec *rec, const EffectParams *P, const EffectState *S, const TypeDev *T,
	uint32_t numGp, float px, float py, float pz, float &ax, float &ay, float &az)
{
	for (uint32_t k = 0; k < numGp; k++) {
		const GravityPointDev &gp = T->gravityPoints[k];
		float wp[3];
#pragma unroll
		for (int i = 0; i < 3; i++) {
			float e3 = P->emitterToWorld[12 + i];
			float pe3 = (rec->flags & 1u) ? S->prevEmitterToWorld[12 + i] : e3;
			float v = e3 + (pe3 - e3);
			v += P->emitterToWorld[0 + i] * gp.position[0];
			v += P->emitterToWorld[4 + i] * gp.position[1];
			v += P->emitterToWorld[8 + i] * gp.position[2];
			wp[i] = v;
		}
		float dx = px - wp[0], dy = py - wp[1], dz = pz - wp[2];
		float lenSq = (dx*dx + dy*dy + dz*dz) + gp.radius;
		float weight = __fdiv_rn(__fdiv_rn(1.0f, __fsqrt_rn(lenSq)), lenSq) * gp.strength;
		ax += dx * weight;
		ay += dy * weight;
		az += dz * weight;
	}
}

// One slot of the integration loop (:552-605). Returns alive (:611) / dead (:556-567); `ra`,`rb`
// hold the post-step record when the slot is not a parked free slot.
template <bool GP>
__device__ __forceinline__ void integrate_slot(const StepConsts &c, float4 &ra, float4 &rb, bool &alive, bool &dead, bool &touched)
{
	float life = ra.w;
	bool wasFree = !(life < kHugeLifeCmp) && life == life; // parked at 1e20: unobservable, skip the arithmetic
	alive = false; dead = false; touched = !wasFree;
	if (life <= 0.0f) { dead = true; life = kHugeLife; }   // :556-567
	if (wasFree) return;
	float px = ra.x, py = ra.y, pz = ra.z;
	float vx = rb.x, vy = rb.y, vz = rb.z, seed = rb.w;
	float ax = c.gx, ay = c.gy, az = c.gz;                  // :574-579
	ax -= vx * c.drag;
	ay -= vy * c.drag;
	az -= vz * c.drag;
	if (GP && c.numGp) gravity_points_accel(c.rec, c.P, c.S, c.T, c.numGp, px, py, pz, ax, ay, az); // :539-550, :581-590 (rare: kept out of line)
	vx += ax * c.dt; vy += ay * c.dt; vz += az * c.dt;      // :592-594
	px += vx * c.dt; py += vy * c.dt; pz += vz * c.dt;      // :598-600
	life -= __fdiv_rn(c.dt, seed * c.lifeVar + c.lifeTime); // :604
	ra = make_float4(px, py, pz, life);
	rb = make_float4(vx, vy, vz, seed);
	alive = life < kHugeLifeCmp;                            // :611
}

__device__ __forceinline__ void step_consts_from(StepConsts &c, const RoundConsts &rc, const TablePtrs &tab, uint32_t g)
{
	c.rec = &tab.plans[rc.planIdx];
	c.P = &tab.params[g];
	c.S = &tab.state[g];
	c.T = &tab.types[rc.typeIdx];
	c.dt = rc.dt; c.drag = rc.drag; c.lifeTime = rc.lifeTime; c.lifeVar = rc.lifeVar;
	c.gx = rc.gx; c.gy = rc.gy; c.gz = rc.gz;
	c.numGp = rc.flags >> 8;
}

// The integration pass of both modes: pure streaming, no barrier and no cross-tile dependency.
// One warp per 128-slot granule: integrate (:552-605), store the state, write the alive / dead
// ballots and -- in sort mode (KEYS) -- the depth key of every live slot. Everything order
// dependent is left to the cheap follow-up kernels that only read the ballots: k_dead_append
// (free-list append, alive ranks, counts) and then either the radix sort (whose first scatter is
// dense by construction, i.e. it IS the compaction) or k_expand_stream (slot-order stream).
// Effects that fit one granule (TINY) are finished right here, inside the warp.
// resident CTAs per SM asked of the compiler: the one-granule (TINY) instantiation at 4 (64 registers, 52 bytes spilled) runs the
// C5 frame's integration 5 % faster than at 3 (80 registers, none spilled): 0.1641 -> 0.1562 ms per launch, tools/runs/r2bf.sh
#ifndef SPR_TINY_MIN_CTAS
#define SPR_TINY_MIN_CTAS 4
#endif
#define SPR_INTEGRATE_MIN_CTAS(GP, TINY) ((GP) ? 3 : (TINY) ? SPR_TINY_MIN_CTAS : 4)
template <bool GP, bool TINY, bool KEYS>
__global__ void __launch_bounds__(256, SPR_INTEGRATE_MIN_CTAS(GP, TINY)) k_integrate_pass(PoolPtrs pool, TablePtrs tab, uint32_t stamp, uint32_t round)
{
	pdl_enter();
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t granule = blockIdx.x * 8 + warp;
	uint32_t g = pool.granuleOwner[granule];
	if (g == kNoOwner) return;
	// Everything that depends only on the owner is requested at once: the effect's round record
	// and the 128 slot records, whose address is a function of the granule alone.
	// Ownership covers only the slots the effect can have reached, so these loads are rarely wasted -- in the first
	// sub-step of a frame. Later rounds exist for the few effects that owe more than one sub-step (C5: one effect in
	// ten): there the round record is looked at first, 80 bytes instead of 4 KB for every effect that is done.
	if (round != 0) {
		const uint32_t rStamp = tab.roundConsts[g].stamp, rRound = tab.roundConsts[g].round;
		if (rStamp != stamp || rRound != round) return;
	}
	float4 a[4], b[4];
	float4 *slotPtr = pool.state + 2ull * ((uint64_t)granule * kGranuleSlots + lane);
#pragma unroll
	for (int r = 0; r < 4; r++) ld_record(slotPtr + r * 64, a[r], b[r]);
	const RoundConsts rc = tab.roundConsts[g];
	if (rc.stamp != stamp || rc.round != round) return;
	const uint32_t slotBase = rc.slotBase;
	const uint32_t localBase = granule * kGranuleSlots - slotBase;
	if (localBase >= rc.slotsPost) return;
	StepConsts c;
	step_consts_from(c, rc, tab, g);
	struct { uint32_t slotsPost; } pe = { rc.slotsPost };

	// running min / max of the live positions (:532, :611-665; the reference also tracks the life lane of its
	// Float4 min / max, but gpuBounds only ever uses xyz, :672-677)
	uint32_t mnx = 0xffffffffu, mny = 0xffffffffu, mnz = 0xffffffffu;
	uint32_t mxx = 0, mxy = 0, mxz = 0;
	uint32_t anyAlive = 0, anyDead = 0, keyMin = 0xffffffffu;
	const uint32_t keyIdx = slotBase + localBase + lane; // indices, not pointers: one live register each
	const uint32_t maskWord = (slotBase + localBase) >> 5;
	const bool tiny = TINY && (rc.flags & 16u); // the whole effect is this one granule: sort and expand in the warp
	uint32_t am[4], dm[4], key[4];
#pragma unroll
	for (int r = 0; r < 4; r++) {
		bool alive = false, dead = false, touched = false;
		float4 ra = a[r], rb = b[r];
		if (localBase + r * 32 + lane < pe.slotsPost) {
			integrate_slot<GP>(c, ra, rb, alive, dead, touched);
			if (touched) st_record(slotPtr + r * 64, ra, rb);
		}
		key[r] = 0;
		if (alive) {
			uint32_t ex = enc_float(ra.x), ey = enc_float(ra.y), ez = enc_float(ra.z);
			mnx = min(mnx, ex); mny = min(mny, ey); mnz = min(mnz, ez);
			mxx = max(mxx, ex); mxy = max(mxy, ey); mxz = max(mxz, ez);
			if (KEYS) {
				key[r] = depth_key(ra.x, ra.y, ra.z, rc.camX, rc.camY, rc.camZ); // sf::lengthSq (src/sf/Vector.h:99), descending
				keyMin = min(keyMin, key[r]);
				if (!tiny) pool.keys[keyIdx + r * 32] = key[r];
			}
		}
		if (TINY) { a[r] = ra; b[r] = rb; }
		am[r] = __ballot_sync(FULL_MASK, alive); dm[r] = __ballot_sync(FULL_MASK, dead);
		if (lane == 0 && !tiny) { pool.aliveMask[maskWord + r] = am[r]; pool.deadMask[maskWord + r] = dm[r]; }
		anyAlive |= am[r];
		anyDead |= dm[r];
	}
	if (tiny) {
		const uint32_t ltMask = (1u << lane) - 1u;
		// free-list append in ascending slot order (:556-567) and the leftover entries (:491-493)
		uint32_t run = rc.topPost;
#pragma unroll
		for (int r = 0; r < 4; r++) {
			if (dm[r] >> lane & 1u) pool.freeStack[slotBase + run + __popc(dm[r] & ltMask)] = r * 32 + lane;
			run += __popc(dm[r]);
		}
		if ((rc.flags & 8u) && lane < rc.topPost) pool.freeStack[slotBase + lane] = rc.slotsPost - 4 + lane;
		// sort mode: stable depth rank by counting, #{j : key_j < key_i, or equal and slot_j < slot_i};
		// stream mode: ascending slot order (:611-665)
		uint32_t rank[4] = { 0, 0, 0, 0 };
		if (!KEYS) {
			uint32_t runA = 0;
#pragma unroll
			for (int r = 0; r < 4; r++) { rank[r] = runA + __popc(am[r] & ltMask); runA += __popc(am[r]); }
		}
#pragma unroll
		for (int rj = 0; KEYS && rj < 4; rj++) {
			uint32_t m = am[rj];
			while (m) {
				uint32_t lj = __ffs(m) - 1;
				m &= m - 1;
				uint32_t kj = __shfl_sync(FULL_MASK, key[rj], lj);
				uint32_t sj = rj * 32 + lj;
#pragma unroll
				for (int r = 0; r < 4; r++) rank[r] += (kj < key[r] || (kj == key[r] && sj < (uint32_t)(r * 32) + lane)) ? 1u : 0u;
			}
		}
		uint32_t total = __popc(am[0]) + __popc(am[1]) + __popc(am[2]) + __popc(am[3]);
#pragma unroll
		for (int r = 0; r < 4; r++) {
			if (am[r] >> lane & 1u) {
				// (writing these through a per-warp staging buffer in transposed order, as k_expand_sorted does, was measured
				// on C5's 100k one-granule effects: no change -- this path is bound by its per-warp latency chain, not by the LSU)
				float4 *d = pool.vertices + 8ull * (slotBase + rank[r]);
				st_record(d, a[r], b[r]); st_record(d + 2, a[r], b[r]); st_record(d + 4, a[r], b[r]); st_record(d + 6, a[r], b[r]);
				if (KEYS) pool.vals[slotBase + rank[r]] = r * 32 + lane;
			}
		}
		if (lane == 0) {
			const uint32_t parIn = rc.flags & 1u;
			c.S->aliveCount = total;                              // :670
			c.S->freeCount[parIn ^ 1u] = run;
			c.S->slotCount[parIn ^ 1u] = rc.slotsPost;
		}
	}
	if (anyDead && lane == 0 && !tiny) tab.effectAnyDead[g] = 1; // idempotent flag, a plain store is enough
	if (anyAlive) {
		mnx = __reduce_min_sync(FULL_MASK, mnx); mny = __reduce_min_sync(FULL_MASK, mny); mnz = __reduce_min_sync(FULL_MASK, mnz);
		mxx = __reduce_max_sync(FULL_MASK, mxx); mxy = __reduce_max_sync(FULL_MASK, mxy); mxz = __reduce_max_sync(FULL_MASK, mxz);
		if (KEYS) keyMin = __reduce_min_sync(FULL_MASK, keyMin);
		if (lane < (KEYS ? 9 : 7) && (lane & 3u) != 3u) { // lanes 0-2: boundsMin, 4-6: boundsMax, 8: keyMin
			uint32_t v = lane == 0 ? mnx : lane == 1 ? mny : lane == 2 ? mnz
				: lane == 4 ? mxx : lane == 5 ? mxy : lane == 6 ? mxz : keyMin;
			uint32_t *addr = &c.S->boundsMin[0] + (lane < 8 ? lane : 8 + warp); // boundsMin[4], boundsMax[4], keyMin[8] are contiguous
			// the current value (read at L2, where the atomics land) filters out almost every atomic: the bounds only ever tighten
			const uint32_t curBound = __ldcg(addr);
			if (lane < 4 || lane == 8) { if (v < curBound) atomicMin(addr, v); }
			else { if (v > curBound) atomicMax(addr, v); }
		}
	}
}

// Single-chain variant of tile_lookback (dead counts only).
__device__ inline uint32_t tile_lookback1(unsigned long long *st, uint32_t poolTile, uint32_t tileIdx, uint32_t agg, uint32_t epoch, uint32_t lane)
{
	const unsigned long long tag = (unsigned long long)epoch << 34;
	const unsigned long long fAgg = 1ull << 32, fIncl = 2ull << 32;
	if (tileIdx == 0) { if (lane == 0) st_status(st + poolTile, tag | fIncl | agg); return 0; }
	if (lane == 0) st_status(st + poolTile, tag | fAgg | agg);
	uint32_t excl = 0;
	int32_t rel = (int32_t)tileIdx - 1 - (int32_t)lane;
	for (;;) {
		unsigned long long v = tag | fIncl;
		if (rel >= 0) { uint32_t t = poolTile - (tileIdx - (uint32_t)rel); uint32_t spins = 0; do { v = ld_status(st + t); spin_guard(spins); } while ((v >> 34) != epoch); }
		uint32_t incl = __ballot_sync(FULL_MASK, ((v >> 32) & 3u) == 2u);
		uint32_t upTo = incl ? (uint32_t)(__ffs(incl) - 1) : 31u;
		excl += __reduce_add_sync(FULL_MASK, lane <= upTo ? (uint32_t)v : 0u);
		if (incl) break;
		rel -= 32;
	}
	if (lane == 0) st_status(st + poolTile, tag | fIncl | (excl + agg));
	return excl;
}

// Everything order dependent, from the ballots alone: the free-list append in ascending slot order
// (:556-567), the effect's new slot / free counts and -- in stream mode (ALIVE) -- the rank of every
// 32-slot word's first live particle in the effect's stream (:611-665) plus gpuParticles.size (:670).
// One warp per 1024-slot pool tile, lane = one mask word; chained look-back across the tiles of a
// large effect. Touches 8 bytes per 32 slots, so the chain is cheap.
template <bool ALIVE>
__global__ void __launch_bounds__(128) k_dead_append(PoolPtrs pool, TablePtrs tab, uint32_t stamp, uint32_t round, uint32_t epoch, uint32_t numPoolTiles,
	unsigned long long ticketBase)
{
	pdl_enter();
	const uint32_t lane = threadIdx.x & 31;
#if SPR_STRICT_TILE_TICKETS
	// the warp's tile = its ticket (every launched warp takes exactly one: the host advances ticketBase by the warp count)
	unsigned long long ticket = 0;
	if (lane == 0) ticket = atomicAdd(pool.tickets + 1, 1ull) - ticketBase;
	const uint32_t poolTile = (uint32_t)__shfl_sync(FULL_MASK, ticket, 0);
#else
	const uint32_t poolTile = blockIdx.x * 4 + (threadIdx.x >> 5); // waits only for lower block indices (device_types.h, SPR_STRICT_TILE_TICKETS)
	(void)ticketBase;
#endif
	if (poolTile >= numPoolTiles) return;
	const uint32_t granule = poolTile * 8 + (lane >> 2);
	uint32_t g = pool.granuleOwner[granule];
	bool active = false;
	uint32_t slotBase = 0, wordBase = 0, word = 0, aword = 0;
	struct { uint32_t slotsPost, topPost, parIn; bool grew; } pe = { 0, 0, 0, false };
	bool large = false, anyDead = false;
	EffectState *S = nullptr;
	if (g != kNoOwner) {
		const RoundConsts rc = tab.roundConsts[g];
		active = rc.stamp == stamp && rc.round == round && !(rc.flags & 16u); // single-granule effects are finished by k_integrate_sort
		if (active) {
			S = &tab.state[g];
			pe.slotsPost = rc.slotsPost; pe.topPost = rc.topPost; pe.parIn = rc.flags & 1u; pe.grew = (rc.flags & 8u) != 0;
			slotBase = rc.slotBase;
			large = (rc.flags & 2u) != 0;
			wordBase = poolTile * kTileSlots + lane * 32 - slotBase; // first local slot of this lane's word
			if (wordBase >= pe.slotsPost) active = false;
			else {
				if (tab.effectAnyDead[g]) { anyDead = true; word = pool.deadMask[(poolTile * kTileSlots >> 5) + lane]; }
				if (ALIVE) aword = pool.aliveMask[(poolTile * kTileSlots >> 5) + lane];
			}
		}
	}
	if (!__any_sync(FULL_MASK, active)) return;
	uint32_t d = __popc(word);
	// a tile wholly inside one large effect (the common case for big effects) scans with log-step shuffles
	const bool uniform = __all_sync(FULL_MASK, large);
	if (ALIVE) { // same two-level scan for the alive ballots; always needed in stream mode
		uint32_t na = __popc(aword), baseA = 0, tileA = 0;
		if (uniform) {
			uint32_t incl = active ? na : 0u;
#pragma unroll
			for (int dlt = 1; dlt < 32; dlt <<= 1) { uint32_t o = __shfl_up_sync(FULL_MASK, incl, dlt); if ((int)lane >= dlt) incl += o; }
			baseA = incl - (active ? na : 0u);
			tileA = __shfl_sync(FULL_MASK, incl, 31);
		} else {
			for (int j = 0; j < 32; j++) {
				uint32_t oj = __shfl_sync(FULL_MASK, active ? g : kNoOwner, j);
				uint32_t aj = __shfl_sync(FULL_MASK, na, j);
				if (active && oj == g) { tileA += aj; if (j < (int)lane) baseA += aj; }
			}
		}
		if (__any_sync(FULL_MASK, large)) {
			uint32_t tileIdx = (poolTile * kTileSlots - slotBase) / kTileSlots;
			baseA += tile_lookback1(pool.tileStatusAlive, poolTile, __shfl_sync(FULL_MASK, tileIdx, 0), __shfl_sync(FULL_MASK, tileA, 0), epoch, lane);
		}
		if (active) {
			pool.wordAliveBase[(poolTile * kTileSlots >> 5) + lane] = baseA;
			if (pe.slotsPost - 1 - wordBase < 32) S->aliveCount = baseA + na;
		}
	}
	// exclusive prefix among the lower lanes of the same effect, and the effect's total inside the tile
	uint32_t base = 0, tileTotal = 0;
	const bool scan = __any_sync(FULL_MASK, anyDead); // nothing died in any effect of this tile: no ranks needed
	if (scan && uniform) {
		uint32_t incl = active ? d : 0u;
#pragma unroll
		for (int dlt = 1; dlt < 32; dlt <<= 1) { uint32_t o = __shfl_up_sync(FULL_MASK, incl, dlt); if ((int)lane >= dlt) incl += o; }
		base = incl - (active ? d : 0u);
		tileTotal = __shfl_sync(FULL_MASK, incl, 31);
	}
	for (int j = 0; scan && !uniform && j < 32; j++) {
		uint32_t oj = __shfl_sync(FULL_MASK, active ? g : kNoOwner, j);
		uint32_t dj = __shfl_sync(FULL_MASK, d, j);
		if (active && oj == g) { tileTotal += dj; if (j < (int)lane) base += dj; }
	}
	if (scan && __any_sync(FULL_MASK, large)) { // a large effect owns the whole tile: all lanes agree
		uint32_t tileIdx = (poolTile * kTileSlots - slotBase) / kTileSlots;
		base += tile_lookback1(pool.tileStatusDead, poolTile, __shfl_sync(FULL_MASK, tileIdx, 0), __shfl_sync(FULL_MASK, tileTotal, 0), epoch, lane);
	}
	if (!active) return;
	uint32_t w = word, rank = base;
	while (w) {
		uint32_t bit = __ffs(w) - 1;
		pool.freeStack[slotBase + pe.topPost + rank] = wordBase + bit;
		rank++;
		w &= w - 1;
	}
	// entries of a partially used new Particle4 stay at the bottom of the stack (:491-493)
	if (wordBase == 0 && pe.grew) {
		for (uint32_t i = 0; i < pe.topPost; i++) pool.freeStack[slotBase + i] = pe.slotsPost - 4 + i;
	}
	// the lane that holds the last slot publishes the new counts
	if (pe.slotsPost - 1 - wordBase < 32) {
		S->freeCount[pe.parIn ^ 1u] = pe.topPost + base + d;
		S->slotCount[pe.parIn ^ 1u] = pe.slotsPost;
	}
}

// ---------------------------------------------------------------------------
// Stream mode, one pass: k_stream_fused
// ---------------------------------------------------------------------------
// The strict drop-in stream (ascending slot order, :611-665) needs, for every live slot, the number
// of live slots before it in the effect -- a scan across the whole effect -- before its 128 bytes of
// vertices can be placed. Doing that inside the streaming pass saves the second read of the record
// (188 B instead of 224 B per particle-step) but puts a chained look-back on the critical path.
// This kernel hides it by software pipelining inside persistent CTAs:
//   iteration i, phase A (tile i): load, integrate, store state, ballots, stage the live records of
//       each warp contiguously in shared memory (two stage buffers), publish the tile's alive / dead
//       aggregates for the chained scan -- nothing in phase A waits on another CTA;
//   iteration i, phase B (tile i-1): resolve that tile's look-back (by now its predecessors, which
//       were processed one iteration ago by the neighbouring CTAs, have published), then expand the
//       staged records 4x with 512-byte warp stores, append the freed slots, publish the counts.
// A ninth warp per CTA does nothing but the chained scan: it resolves tile i-1 while the eight
// worker warps run phase A of tile i, so the look-back latency never stalls a barrier.
// Persistent grid = resident CTAs only, so every predecessor a CTA waits on is running.

struct FusedTileMeta {        // per stage buffer, written in phase A, read in phase B
	uint32_t owner[8];        // effect of each warp's granule (kNoOwner when the warp had nothing to do)
	uint32_t nAlive[8], nDead[8];
	uint32_t deadMask[8][4];
	uint32_t localBase[8];
	uint32_t slotBase[8], slotsPost[8], topPost[8], flags[8];
	uint32_t tile;            // pool tile index, 0xffffffff when the buffer is empty
	uint32_t large;           // the tile belongs to one large effect
	uint32_t exclA, exclD;    // resolved look-back
};

__device__ inline void tile_lookback2(unsigned long long *stA, unsigned long long *stD, uint32_t poolTile, uint32_t tileIdx,
	uint32_t aggA, uint32_t aggD, uint32_t epoch, uint32_t lane, uint32_t &exclA, uint32_t &exclD)
{
	// the aggregates were already published in phase A; resolve both chains independently
	const unsigned long long tag = (unsigned long long)epoch << 34;
	const unsigned long long fIncl = 2ull << 32;
	exclA = 0; exclD = 0;
	if (tileIdx != 0) {
		int32_t rel = (int32_t)tileIdx - 1 - (int32_t)lane;
		bool doneA = false, doneD = false;
		while (!doneA || !doneD) {
			unsigned long long va = tag | fIncl, vd = tag | fIncl; // virtual zero prefix in front of tile 0
			if (rel >= 0) {
				uint32_t t = poolTile - (tileIdx - (uint32_t)rel);
				uint32_t spins = 0;
				if (!doneA) do { va = ld_status(stA + t); spin_guard(spins); } while ((va >> 34) != epoch);
				if (!doneD) do { vd = ld_status(stD + t); spin_guard(spins); } while ((vd >> 34) != epoch);
			}
			if (!doneA) {
				uint32_t incl = __ballot_sync(FULL_MASK, ((va >> 32) & 3u) == 2u);
				uint32_t upTo = incl ? (uint32_t)(__ffs(incl) - 1) : 31u;
				exclA += __reduce_add_sync(FULL_MASK, lane <= upTo ? (uint32_t)va : 0u);
				doneA = incl != 0;
			}
			if (!doneD) {
				uint32_t incl = __ballot_sync(FULL_MASK, ((vd >> 32) & 3u) == 2u);
				uint32_t upTo = incl ? (uint32_t)(__ffs(incl) - 1) : 31u;
				exclD += __reduce_add_sync(FULL_MASK, lane <= upTo ? (uint32_t)vd : 0u);
				doneD = incl != 0;
			}
			rel -= 32;
		}
	}
	if (lane == 0) { st_status(stA + poolTile, tag | fIncl | (exclA + aggA)); st_status(stD + poolTile, tag | fIncl | (exclD + aggD)); }
}

// Bulk asynchronous stores (TMA): the vertex stream seen as a tensor [particle][copy 0..3][8 floats]; one
// cp.async.bulk.tensor.3d with box [128][1][8] writes copy c of 128 consecutive particles out of 4 KB of compact
// records in shared memory -- four of them expand a full granule without a single LSU store (SASS: UTMASTG.3D).
__device__ __forceinline__ void tma_store_granule(const CUtensorMap *map, const void *stage, uint32_t firstParticle)
{
	const uint32_t s = (uint32_t)__cvta_generic_to_shared(stage);
#pragma unroll
	for (int c = 0; c < 4; c++)
		asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
			:: "l"((uint64_t)map), "r"(s), "r"(0), "r"(c), "r"((int)firstParticle) : "memory");
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_wait_stage_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// The record prefetch is speculative in the first sub-step of a frame (every owned granule is almost certainly
// simulated); later rounds exist for the few effects that owe more than one sub-step (C5: one in ten), so there the
// round record is looked at first -- 8 bytes instead of 4 KB for every effect that is done. (Phase A repeats the test on
// the full record and never touches the registers of a granule that is not active.)
__device__ __forceinline__ bool round_wants(const TablePtrs &tab, uint32_t g, uint32_t stamp, uint32_t round)
{
	if (round == 0) return true;
	return tab.roundConsts[g].stamp == stamp && tab.roundConsts[g].round == round;
}

// The next pool tile with an owner, by ticket. Tiles nobody owns (the tail of every power-of-two slot range behind the
// effect's slot bound: 63 of 128 tiles in a 65536-particle effect) are consumed right here: as iterations of their
// own they cost nothing, but they let CTAs run ahead of the ones working on real tiles, and a CTA that is early at its
// next real tile then waits in that tile's look-back for predecessors that were claimed earlier and are still a full
// iteration away (measured on 128 x 65536 particles: 80 % of the stall samples on the barrier behind the look-back,
// 1.08 ms per step against 0.35 ms for one 10M-particle effect).
__device__ __forceinline__ uint32_t claim_owned_tile(const PoolPtrs &pool, uint32_t numPoolTiles, unsigned long long ticketBase)
{
	for (;;) {
		const unsigned long long t = atomicAdd(pool.tickets + 0, 1ull) - ticketBase;
		if (t >= numPoolTiles) return 0xffffffffu;
		const uint4 *own = (const uint4*)(pool.granuleOwner + (uint32_t)t * 8u); // 8 granules per tile, 32-byte aligned
		const uint4 x = __ldg(own), y = __ldg(own + 1);
		if ((x.x & x.y & x.z & x.w & y.x & y.y & y.z & y.w) != kNoOwner) return (uint32_t)t;
	}
}

template <bool GP, bool TMA>
__global__ void __launch_bounds__(288, 3) k_stream_fused(PoolPtrs pool, TablePtrs tab, uint32_t stamp, uint32_t round, uint32_t epoch, uint32_t numPoolTiles,
	unsigned long long ticketBase, const __grid_constant__ CUtensorMap vertexMap)
{
	pdl_enter();
	extern __shared__ __align__(128) float4 sStageAll[]; // [2][8 warps][256 float4] = 64 KB
	__shared__ FusedTileMeta sMeta[2];
	__shared__ uint32_t sTicket[2];                    // tile of iteration it (parity it & 1), claimed one iteration ahead
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t ltMask = (1u << lane) - 1u;
	const unsigned long long tag = (unsigned long long)epoch << 34;

	// Tiles are claimed with an atomic ticket, in the order the CTAs get to them: every tile below a claimed one belongs
	// to a CTA that is running, so the chained look-back cannot wait for a CTA that was never scheduled. A CTA claims
	// until its ticket runs past the last tile: numPoolTiles + gridDim tickets per launch (the host's ticketBase).
	if (threadIdx.x < 2) sMeta[threadIdx.x].tile = 0xffffffffu;
	if (threadIdx.x == 0) sTicket[0] = claim_owned_tile(pool, numPoolTiles, ticketBase);
	__syncthreads();

	// the records of the NEXT tile are requested before phase B of the current iteration, so their
	// DRAM latency hides behind the expansion stores
	float4 a[4], b[4];
	uint32_t gNext = kNoOwner;
	if (warp < 8 && sTicket[0] < numPoolTiles) {
		const uint32_t first = sTicket[0];
		gNext = pool.granuleOwner[first * 8 + warp];
		if (gNext != kNoOwner && round_wants(tab, gNext, stamp, round)) {
			const float4 *p = pool.state + 2ull * ((uint64_t)(first * 8 + warp) * kGranuleSlots + lane);
#pragma unroll
			for (int r = 0; r < 4; r++) ld_record(p + r * 64, a[r], b[r]);
		}
	}

	for (uint32_t it = 0; ; it++) {
		const uint32_t cur = it & 1u;
		const uint32_t tile = sTicket[cur];            // 0xffffffff once the tickets have run out
		FusedTileMeta &M = sMeta[cur];
		FusedTileMeta &Q = sMeta[cur ^ 1u];
		// claim the tile of the next iteration now: it is read (for the prefetch) behind the barrier that ends phase A
		// (claimed by the scan warp after its look-back instead, the claim delays that barrier: measured 0.32 -> 0.89 ms
		// on 128 x 65536 particles, where every other claim walks over a run of unowned tiles)
		if (threadIdx.x == 0) sTicket[cur ^ 1u] = tile < numPoolTiles ? claim_owned_tile(pool, numPoolTiles, ticketBase) : 0xffffffffu;
		// ---------------- scan warp: look-back of the tile staged one iteration ago, overlapping phase A
		if (warp == 8) {
			if (it > 0 && Q.tile != 0xffffffffu && Q.large) {
				uint32_t tA = 0, tD = 0;
				for (uint32_t w = 0; w < 8; w++) if (Q.owner[w] != kNoOwner) { tA += Q.nAlive[w]; tD += Q.nDead[w]; }
				uint32_t first = 0;
				while (Q.owner[first] == kNoOwner) first++;
				uint32_t tileIdx = (Q.tile * kTileSlots - Q.slotBase[first]) / kTileSlots;
				uint32_t eA, eD;
				tile_lookback2(pool.tileStatusAlive, pool.tileStatusDead, Q.tile, tileIdx, tA, tD, epoch, lane, eA, eD);
				if (lane == 0) { Q.exclA = eA; Q.exclD = eD; }
			}
		}
		// ---------------- phase A (worker warps): tile `tile` into stage buffer `cur`
		if (tile < numPoolTiles && warp < 8) {
			const uint32_t granule = tile * 8 + warp;
			const uint32_t g = gNext;
			bool active = false;
			uint32_t nA = 0, nD = 0, dmask[4] = { 0, 0, 0, 0 };
			RoundConsts rc;
			rc.slotBase = 0; rc.slotsPost = 0; rc.topPost = 0; rc.flags = 0;
			uint32_t localBase = 0;
			if (g != kNoOwner) {
				float4 *slotPtr = pool.state + 2ull * ((uint64_t)granule * kGranuleSlots + lane);
				rc = tab.roundConsts[g];
				localBase = granule * kGranuleSlots - rc.slotBase;
				active = rc.stamp == stamp && rc.round == round && localBase < rc.slotsPost;
				if (active) {
					StepConsts c;
					step_consts_from(c, rc, tab, g);
					uint32_t mnx = 0xffffffffu, mny = 0xffffffffu, mnz = 0xffffffffu;
					uint32_t mxx = 0, mxy = 0, mxz = 0;
					float4 *stage = sStageAll + (cur * 8 + warp) * 256;
					if (TMA) {
						// this warp's bulk stores of the previous iteration (its phase B) read this buffer; lane 0 issued them
						if (lane == 0) tma_wait_stage_read();
						__syncwarp();
					}
#pragma unroll
					for (int r = 0; r < 4; r++) {
						bool alive = false, dead = false, touched = false;
						float4 ra = a[r], rb = b[r];
						if (localBase + r * 32 + lane < rc.slotsPost) {
							integrate_slot<GP>(c, ra, rb, alive, dead, touched);
							if (touched) st_record(slotPtr + r * 64, ra, rb);
						}
						uint32_t am = __ballot_sync(FULL_MASK, alive);
						dmask[r] = __ballot_sync(FULL_MASK, dead);
						if (alive) {
							uint32_t ex = enc_float(ra.x), ey = enc_float(ra.y), ez = enc_float(ra.z);
							mnx = min(mnx, ex); mny = min(mny, ey); mnz = min(mnz, ez);
							mxx = max(mxx, ex); mxy = max(mxy, ey); mxz = max(mxz, ez);
							uint32_t k = nA + __popc(am & ltMask);
							stage[2 * k] = ra;
							stage[2 * k + 1] = rb;
						}
						nA += __popc(am);
						nD += __popc(dmask[r]);
					}
					if (nA) {
						mnx = __reduce_min_sync(FULL_MASK, mnx); mny = __reduce_min_sync(FULL_MASK, mny); mnz = __reduce_min_sync(FULL_MASK, mnz);
						mxx = __reduce_max_sync(FULL_MASK, mxx); mxy = __reduce_max_sync(FULL_MASK, mxy); mxz = __reduce_max_sync(FULL_MASK, mxz);
						if (lane < 7 && lane != 3) {
							uint32_t v = lane == 0 ? mnx : lane == 1 ? mny : lane == 2 ? mnz
								: lane == 4 ? mxx : lane == 5 ? mxy : mxz;
							uint32_t *addr = lane < 4 ? &tab.state[g].boundsMin[lane] : &tab.state[g].boundsMax[lane - 4];
							const uint32_t curBound = __ldcg(addr); // read late and at L2: filters out almost every atomic (the bounds only tighten)
							if (lane < 4) { if (v < curBound) atomicMin(addr, v); }
							else { if (v > curBound) atomicMax(addr, v); }
						}
					}
				}
			}
			if (lane == 0) {
				M.owner[warp] = active ? g : kNoOwner;
				M.nAlive[warp] = nA; M.nDead[warp] = nD;
				M.localBase[warp] = localBase;
				M.slotBase[warp] = rc.slotBase; M.slotsPost[warp] = rc.slotsPost; M.topPost[warp] = rc.topPost; M.flags[warp] = rc.flags;
			}
			if (lane < 4) M.deadMask[warp][lane] = lane == 0 ? dmask[0] : lane == 1 ? dmask[1] : lane == 2 ? dmask[2] : dmask[3];
			if (TMA) fence_async_shared(); // the staged records become visible to the async proxy (phase B's bulk stores)
		}
		__syncthreads();
		if (tile < numPoolTiles && warp == 8) {
			// publish the tile's aggregates (large effects only: small ones never cross a tile)
			bool anyActive = false, large = false;
			uint32_t tA = 0, tD = 0;
			for (uint32_t w = 0; w < 8; w++) {
				if (M.owner[w] != kNoOwner) { anyActive = true; large = (M.flags[w] & 2u) != 0; tA += M.nAlive[w]; tD += M.nDead[w]; }
			}
			if (lane == 0) {
				M.tile = anyActive ? tile : 0xffffffffu;
				M.large = large ? 1u : 0u;
				if (anyActive && large) {
					st_status(pool.tileStatusAlive + tile, tag | (1ull << 32) | tA);
					st_status(pool.tileStatusDead + tile, tag | (1ull << 32) | tD);
				}
			}
		} else if (tile >= numPoolTiles && threadIdx.x == 0) {
			M.tile = 0xffffffffu;
		}

		// ---------------- prefetch (worker warps): owner and records of this CTA's next tile
		if (warp < 8) {
			const uint32_t nextTile = sTicket[cur ^ 1u];
			gNext = kNoOwner;
			if (tile < numPoolTiles && nextTile < numPoolTiles) {
				gNext = pool.granuleOwner[nextTile * 8 + warp];
				if (gNext != kNoOwner && round_wants(tab, gNext, stamp, round)) {
					const float4 *p = pool.state + 2ull * ((uint64_t)(nextTile * 8 + warp) * kGranuleSlots + lane);
#pragma unroll
					for (int r = 0; r < 4; r++) ld_record(p + r * 64, a[r], b[r]);
				}
			}
		}
		// ---------------- phase B (worker warps): the tile staged one iteration ago (buffer cur ^ 1)
		if (it > 0 && warp < 8 && Q.tile != 0xffffffffu && Q.owner[warp] != kNoOwner) {
			const uint32_t g = Q.owner[warp];
			uint32_t baseA = Q.large ? Q.exclA : 0u, baseD = Q.large ? Q.exclD : 0u;
			for (uint32_t w = 0; w < warp; w++) if (Q.owner[w] == g) { baseA += Q.nAlive[w]; baseD += Q.nDead[w]; }
			const uint32_t nA = Q.nAlive[warp], nD = Q.nDead[warp];
			const uint32_t slotBase = Q.slotBase[warp], slotsPost = Q.slotsPost[warp], topPost = Q.topPost[warp];
			const uint32_t localBase = Q.localBase[warp], flags = Q.flags[warp];
			// free-list append in ascending slot order (:556-567), behind the post-emission stack
			if (nD) {
				uint32_t run = baseD;
#pragma unroll
				for (int r = 0; r < 4; r++) {
					uint32_t dm = Q.deadMask[warp][r];
					if (dm >> lane & 1u) pool.freeStack[slotBase + topPost + run + __popc(dm & ltMask)] = localBase + r * 32 + lane;
					run += __popc(dm);
				}
			}
			if (localBase == 0 && (flags & 8u) && lane < topPost) pool.freeStack[slotBase + lane] = slotsPost - 4 + lane; // :491-493
			// 4x expansion of the staged run with fully coalesced 512-byte warp stores (:611-665)
			const float4 *stage = sStageAll + ((cur ^ 1u) * 8 + warp) * 256;
			float4 *dst = pool.vertices + 8ull * (slotBase + baseA);
			const uint32_t total = nA * 8;
			if (TMA && nA == kGranuleSlots) {
				if (lane == 0) tma_store_granule(&vertexMap, stage, slotBase + baseA); // a full granule: 4 bulk stores of 4 KB
			} else
			for (uint32_t q = lane; q < total; q += 32) dst[q] = stage[((q >> 3) << 1) | (q & 1u)];
			// the warp that holds the last slot publishes the effect's new counts (:670)
			if (lane == 0 && slotsPost - 1 - localBase < kGranuleSlots) {
				EffectState &S = tab.state[g];
				S.aliveCount = baseA + nA;
				S.freeCount[(flags & 1u) ^ 1u] = topPost + baseD + nD;
				S.slotCount[(flags & 1u) ^ 1u] = slotsPost;
			}
		}
		if (tile >= numPoolTiles) break;
		__syncthreads(); // the buffer just drained is the next iteration's phase-A target
	}
	if (TMA && warp < 8 && lane == 0) tma_wait_stage_read(); // shared memory stays allocated until the bulk stores have read it
}

// Stream mode: the 4x vertex expansion in ascending slot order (:607-665). One warp per granule,
// no dependency on any other warp: the rank of a word's first live particle comes from
// k_dead_append, the rest is a popcount. A lane re-reads its 32-byte record (256-bit, streaming) and
// stores it four times; for a dense population the 32 lanes fill 4 KB of contiguous vertex stream.
__global__ void __launch_bounds__(256) k_expand_stream(PoolPtrs pool, TablePtrs tab, uint32_t stamp)
{
	pdl_enter();
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t granule = blockIdx.x * 8 + warp;
	uint32_t g = pool.granuleOwner[granule];
	if (g == kNoOwner) return;
	const uint32_t word0 = granule * (kGranuleSlots / 32);
	float4 a[4], b[4];
	const float4 *slotPtr = pool.state + 2ull * ((uint64_t)granule * kGranuleSlots + lane);
#pragma unroll
	for (int r = 0; r < 4; r++) ld_record(slotPtr + r * 64, a[r], b[r]);
	uint32_t am = lane < 4 ? pool.aliveMask[word0 + lane] : 0u;
	uint32_t base = lane < 4 ? pool.wordAliveBase[word0 + lane] : 0u;
	const RoundConsts rc = tab.roundConsts[g];
	if (rc.stamp != stamp || (rc.flags & 16u)) return; // not simulated this frame, or finished inside k_integrate_pass
	const uint32_t localBase = granule * kGranuleSlots - rc.slotBase;
	if (localBase >= rc.slotsPost) return;
	float4 *dst = pool.vertices + 8ull * rc.slotBase;
#pragma unroll
	for (int r = 0; r < 4; r++) {
		uint32_t m = __shfl_sync(FULL_MASK, am, r), bs = __shfl_sync(FULL_MASK, base, r);
		if (localBase + r * 32 >= rc.slotsPost) m = 0; // words past the end were not rewritten this step
		if (m >> lane & 1u) {
			float4 *d = dst + 8ull * (bs + __popc(m & ((1u << lane) - 1u)));
			st_record(d, a[r], b[r]); st_record(d + 2, a[r], b[r]); st_record(d + 4, a[r], b[r]); st_record(d + 6, a[r], b[r]);
		}
	}
}

// prevEmitterToWorld = emitterToWorld after the frame's sub-steps (:668)
__global__ void k_finalize(TablePtrs tab, const ActiveEntry *active, uint32_t numActive, uint32_t stamp)
{
	pdl_launch_dependents();
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t a = t >> 4, i = t & 15;
	if (a >= numActive) return;
	ActiveEntry ae = active[a];               // host-uploaded entry list: read ahead of the wait
	if (ae.numSubsteps == 0) return;
	pdl_wait();
	tab.state[ae.effect].prevEmitterToWorld[i] = tab.params[ae.effect].emitterToWorld[i];
	// the exact slot count after this frame, for the host's capacity bound (read there without synchronising): only for
	// effects whose bound is not final -- the mirror lives in pinned host memory, 8 bytes over PCIe per effect and frame
	if (i == 0 && (ae.flags & 2u)) {
		const EffectState &S = tab.state[ae.effect];
		tab.slotMirror[ae.effect] = (unsigned long long)stamp << 32 | S.slotCount[S.parity];
	}
}

// ---------------------------------------------------------------------------
// Render-side gather: what renderMain reads per visible effect (:838-846)
// ---------------------------------------------------------------------------

__device__ inline float dec_float(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void k_gather_render(TablePtrs tab, const uint32_t *effects, uint32_t n, RenderGather *out)
{
	pdl_enter();
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const EffectState &S = tab.state[effects[i]];
	RenderGather r;
	r.aliveCount = S.aliveCount;
	r.slotCount = S.slotCount[S.parity];
	r.freeCount = S.freeCount[S.parity];
	r.stopEmit = S.stopEmit;
	r.firstEmit = S.firstEmit;
	r.spawnTimer = S.spawnTimer;
	r.emitOnTimer = S.emitOnTimer;
#pragma unroll
	for (int k = 0; k < 4; k++) { r.boundsMin[k] = dec_float(S.boundsMin[k]); r.boundsMax[k] = dec_float(S.boundsMax[k]); }
	out[i] = r;
}

// ---------------------------------------------------------------------------
// Area query: which effects' sphere areas intersect a frustum
// (AreaSystem::queryFrustum, src/client/AreaSystem.cpp:139-170, over the spheres the particle system registers,
// ParticleSystem.cpp:762-763, :782-783; sf::Frustum::intersects(Sphere), src/sf/Frustum.h:42-58)
// ---------------------------------------------------------------------------

struct FrustumDev { float side[4][4], caps[4][4]; }; // SprFrustum: side[k][j] = component k of plane j

__global__ void __launch_bounds__(256) k_query_frustum(TablePtrs tab, uint32_t numGlobals, uint32_t systemIdx, FrustumDev f,
	uint32_t *outIds, uint32_t capacity, uint32_t *outCount)
{
	pdl_enter();
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
	bool hit = false;
	uint32_t id = 0;
	if (g < numGlobals) {
		const EffectParams &P = tab.params[g];
		if (P.live && P.systemIdx == systemIdx) {
			id = P.effectId;
			const float ox = P.areaSphere[0], oy = P.areaSphere[1], oz = P.areaSphere[2], r = P.areaSphere[3];
			hit = true;
#pragma unroll
			for (int j = 0; j < 4; j++) {
				// so = side[3] + r; so += side[0] * o.x; so += side[1] * o.y; so += side[2] * o.z (Frustum.h:44-55), same for caps
				float so = f.side[3][j] + r, co = f.caps[3][j] + r;
				so += f.side[0][j] * ox; co += f.caps[0][j] * ox;
				so += f.side[1][j] * oy; co += f.caps[1][j] * oy;
				so += f.side[2][j] * oz; co += f.caps[2][j] * oz;
				hit = hit && (so < co ? so : co) > 0.0f; // so.min(co).allGreaterThanZero(), :57
			}
		}
	}
	const uint32_t ballot = __ballot_sync(FULL_MASK, hit);
	if (!ballot) return;
	uint32_t base = 0;
	if (lane == 0) base = atomicAdd(outCount, __popc(ballot));
	base = __shfl_sync(FULL_MASK, base, 0);
	const uint32_t pos = base + __popc(ballot & ((1u << lane) - 1u));
	if (hit && pos < capacity) outIds[pos] = id;
}

// ---------------------------------------------------------------------------
// Multi-GPU run packing / expansion
// ---------------------------------------------------------------------------

// offsets[i] = sum of aliveCount of effects[0..i), offsets[n] = total (single CTA; n is small)
__global__ void k_run_offsets(TablePtrs tab, const uint32_t *effects, uint32_t n, uint64_t *offsets, uint32_t *counts)
{
	pdl_enter();
	__shared__ uint64_t sWarp[32];
	__shared__ uint64_t sCarry;
	if (threadIdx.x == 0) sCarry = 0;
	__syncthreads();
	for (uint32_t base = 0; base < n; base += blockDim.x) {
		uint32_t i = base + threadIdx.x;
		uint64_t v = i < n ? tab.state[effects[i]].aliveCount : 0;
		if (i < n && counts) counts[i] = (uint32_t)v;
		uint64_t incl = v;
		for (int d = 1; d < 32; d <<= 1) { uint64_t o = __shfl_up_sync(FULL_MASK, incl, d); if ((threadIdx.x & 31) >= d) incl += o; }
		if ((threadIdx.x & 31) == 31) sWarp[threadIdx.x >> 5] = incl;
		__syncthreads();
		uint64_t warpOff = 0;
		for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) warpOff += sWarp[w];
		uint64_t carry = sCarry;
		if (i < n) offsets[i] = carry + warpOff + incl - v;
		__syncthreads();
		if (threadIdx.x == blockDim.x - 1) sCarry = carry + warpOff + incl;
		__syncthreads();
	}
	if (threadIdx.x == 0) offsets[n] = sCarry;
}

// Half `half` of record k of an effect's run. FROM_STATE = false: every 4th vertex of the effect's stream;
// true (deferred expansion: the local stream was not written): the slot record behind the k-th sorted slot.
template <bool FROM_STATE>
__device__ __forceinline__ float4 run_record_half(const PoolPtrs &pool, uint32_t slotBase, uint64_t k, uint32_t half)
{
	if (FROM_STATE) return pool.state[2ull * (slotBase + pool.vals[slotBase + k]) + half];
	return pool.vertices[8ull * (slotBase + k) + half];
}

// out float4 q -> record q>>1 of the concatenated runs
template <bool FROM_STATE>
__global__ void __launch_bounds__(256) k_pack_runs(PoolPtrs pool, TablePtrs tab, const uint32_t *effects, uint32_t n,
	const uint64_t *offsets, float4 *out, uint64_t capacityRecords)
{
	pdl_enter();
	uint64_t total = offsets[n];
	if (total > capacityRecords) total = capacityRecords;
	uint64_t totalQ = total * 2;
	for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < totalQ; q += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t rcd = q >> 1;
		uint32_t lo = 0, hi = n;
		while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (offsets[mid] <= rcd) lo = mid; else hi = mid; }
		uint64_t k = rcd - offsets[lo];
		uint32_t slotBase = tab.params[effects[lo]].slotBase;
		out[q] = run_record_half<FROM_STATE>(pool, slotBase, k, (uint32_t)(q & 1u));
	}
}

// 4x expansion of gathered 32-byte records into the vertex layout (:611-665)
__global__ void __launch_bounds__(256) k_expand_runs(const float4 *records, uint64_t numRecords, float4 *out)
{
	pdl_enter();
	uint64_t totalQ = numRecords * 8;
	for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < totalQ; q += (uint64_t)gridDim.x * blockDim.x) {
		out[q] = records[((q >> 3) << 1) | (q & 1u)];
	}
}

// Run buffer layout: 32-byte header {uint64 count, pad} followed by `count` 32-byte records
// (32-byte aligned, so that a record moves with one 256-bit access).
// Same as k_pack_runs, but the record count goes into the header of the destination buffer.
template <bool FROM_STATE>
__global__ void __launch_bounds__(256) k_pack_runs_buf(PoolPtrs pool, TablePtrs tab, const uint32_t *effects, uint32_t n,
	const uint64_t *offsets, float4 *buf, uint64_t capacityRecords)
{
	pdl_enter();
	uint64_t total = offsets[n];
	if (total > capacityRecords) total = capacityRecords;
	if (blockIdx.x == 0 && threadIdx.x == 0) ((unsigned long long*)buf)[0] = total;
	float4 *out = buf + 2;
	uint64_t totalQ = total * 2;
	for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < totalQ; q += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t rcd = q >> 1;
		uint32_t lo = 0, hi = n;
		while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (offsets[mid] <= rcd) lo = mid; else hi = mid; }
		uint64_t k = rcd - offsets[lo];
		uint32_t slotBase = tab.params[effects[lo]].slotBase;
		out[q] = run_record_half<FROM_STATE>(pool, slotBase, k, (uint32_t)(q & 1u));
	}
}

// Fused all-gather + 4x expansion: every rank's run is read straight from (peer) memory over
// NVLink with plain 256-bit loads on the IPC-mapped pointers and written four times into the local
// draw stream (:607-665 layout), rank-major. One lane per record; the loads of a warp cover 1 KB of
// contiguous peer memory, its stores 4 KB of contiguous local memory.
struct PeerRuns { const float4 *buf[16]; uint32_t numRanks; uint32_t myRank; };
__global__ void __launch_bounds__(256) k_expand_from_peers(PeerRuns runs, float4 *out, uint64_t capacityRecords, unsigned long long *totalOut)
{
	pdl_enter();
	__shared__ unsigned long long sOffset[17];
	__shared__ unsigned long long sMaxChunks;
	if (threadIdx.x == 0) {
		unsigned long long run = 0, mx = 0;
		for (uint32_t r = 0; r < runs.numRanks; r++) {
			sOffset[r] = run;
			unsigned long long c;
			asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(c) : "l"(runs.buf[r]) : "memory");
			if (run + c > capacityRecords) c = capacityRecords > run ? capacityRecords - run : 0;
			run += c;
			unsigned long long chunks = (c + 1023) / 1024;
			if (chunks > mx) mx = chunks;
		}
		sOffset[runs.numRanks] = run;
		sMaxChunks = mx;
		if (blockIdx.x == 0 && totalOut) *totalOut = run;
	}
	__syncthreads();
	// Work item j -> 1024-record chunk (j / numRanks) of rank (j + myRank) % numRanks: at any moment the
	// CTAs of one GPU read from ALL peers, and the GPUs start on different peers, so no source GPU's
	// NVLink egress is shared by every reader at once (the rank-major order was 3x slower at 8 GPUs).
	const unsigned long long items = sMaxChunks * runs.numRanks;
	for (unsigned long long j = blockIdx.x; j < items; j += gridDim.x) {
		const uint32_t r = (uint32_t)((j + runs.myRank) % runs.numRanks);
		const unsigned long long count = sOffset[r + 1] - sOffset[r];
		const unsigned long long base = (j / runs.numRanks) * 1024;
		if (base >= count) continue;
		const float4 *src = runs.buf[r] + 2 + 2ull * base;
		float4 *dst = out + 8ull * (sOffset[r] + base);
		float4 a[4], b[4];
		bool ok[4];
#pragma unroll
		for (int u = 0; u < 4; u++) {
			a[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); b[u] = a[u];
			unsigned long long i = base + u * 256 + threadIdx.x;
			ok[u] = i < count;
			if (ok[u]) {
				asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
					: "=f"(a[u].x), "=f"(a[u].y), "=f"(a[u].z), "=f"(a[u].w), "=f"(b[u].x), "=f"(b[u].y), "=f"(b[u].z), "=f"(b[u].w)
					: "l"(src + 2ull * (u * 256 + threadIdx.x)));
			}
		}
		// transposed across the warp like k_expand_sorted: every store instruction covers 1 KB of consecutive bytes
		const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
		for (int u = 0; u < 4; u++) {
			const unsigned long long wbase = base + u * 256 + warp * 32;
			if (wbase >= count) continue; // warp-uniform
			const uint32_t cnt = count - wbase < 32 ? (uint32_t)(count - wbase) : 32u;
			float4 *d = dst + 8ull * (u * 256 + warp * 32);
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const uint32_t p = k * 8 + (lane >> 2);
				float4 x, y;
				x.x = __shfl_sync(FULL_MASK, a[u].x, p); x.y = __shfl_sync(FULL_MASK, a[u].y, p); x.z = __shfl_sync(FULL_MASK, a[u].z, p); x.w = __shfl_sync(FULL_MASK, a[u].w, p);
				y.x = __shfl_sync(FULL_MASK, b[u].x, p); y.y = __shfl_sync(FULL_MASK, b[u].y, p); y.z = __shfl_sync(FULL_MASK, b[u].z, p); y.w = __shfl_sync(FULL_MASK, b[u].w, p);
				if (p < cnt) st_record(d + 2 * (k * 32 + lane), x, y);
			}
		}
	}
}

// The same assembly with NO register held per byte in flight, for the overlapped pipeline where the kernel is confined to
// a small share of every SM (sprContextSetAssemblyShare): lane 0 of every warp drives a pipeline of kPeerStages bulk
// asynchronous loads (cp.async.bulk global -> shared, 4 KB = 128 records each, completion on an mbarrier) out of the
// peers' runs and hands every landed chunk to four bulk tensor stores (one per copy of the quad, the vertex stream as
// [record][copy][8 floats]) -- neither the pull nor the 4x expansion issues an LSU instruction on the data. Runs are cut
// into full 128-record chunks; the tail of every run (< 128 records) goes through the warp's lanes at the end.
static const int kPeerWarps = 4, kPeerStages = 3;       // 48 KB of shared memory per CTA: fits beside four k_sort_pass CTAs
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(kPeerWarps * 32) k_expand_from_peers_tma(PeerRuns runs, float4 *out, uint64_t capacityRecords, unsigned long long *totalOut,
	const __grid_constant__ CUtensorMap map)
{
	pdl_enter();
	extern __shared__ __align__(128) uint8_t sPeerStage[];
	__shared__ __align__(8) unsigned long long sBar[kPeerWarps][kPeerStages];
	__shared__ unsigned long long sOffset[17];
	__shared__ unsigned long long sMaxFull;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) {
		unsigned long long run = 0, mx = 0;
		for (uint32_t r = 0; r < runs.numRanks; r++) {
			sOffset[r] = run;
			unsigned long long c;
			asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(c) : "l"(runs.buf[r]) : "memory");
			if (run + c > capacityRecords) c = capacityRecords > run ? capacityRecords - run : 0;
			run += c;
			if (c / 128 > mx) mx = c / 128;
		}
		sOffset[runs.numRanks] = run;
		sMaxFull = mx;
		if (blockIdx.x == 0 && totalOut) *totalOut = run;
	}
	if (lane == 0) {
		for (int s = 0; s < kPeerStages; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(&sBar[warp][s])) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	const uint64_t gw = (uint64_t)blockIdx.x * kPeerWarps + warp, stride = (uint64_t)gridDim.x * kPeerWarps;
	if (lane == 0) {
		// work item j -> full chunk (j / numRanks) of rank (j + myRank) % numRanks, as in k_expand_from_peers: every GPU reads all
		// peers at any moment and starts on a different one
		const uint64_t items = sMaxFull * runs.numRanks;
		uint8_t *stage = sPeerStage + (size_t)warp * kPeerStages * 4096;
		auto valid = [&](uint64_t j, uint32_t &r, uint64_t &c) {
			r = (uint32_t)((j + runs.myRank) % runs.numRanks); c = j / runs.numRanks;
			return c < (sOffset[r + 1] - sOffset[r]) / 128;
		};
		auto issue_load = [&](int s, uint32_t r, uint64_t c) {
			const uint32_t bar = smem_addr(&sBar[warp][s]);
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(4096) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				:: "r"(smem_addr(stage + s * 4096)), "l"(runs.buf[r] + 2 + 256ull * c), "r"(4096), "r"(bar) : "memory");
		};
		uint64_t issueJ = gw;
		uint32_t r; uint64_t c;
		uint32_t qr[kPeerStages]; uint64_t qc[kPeerStages]; // what sits in each stage
		int filled = 0;
		for (; filled < kPeerStages && issueJ < items; issueJ += stride) {
			if (!valid(issueJ, r, c)) continue;
			qr[filled] = r; qc[filled] = c;
			issue_load(filled, r, c);
			filled++;
		}
		uint32_t phase = 0;
		int s = 0;
		while (filled > 0) {
			const uint32_t bar = smem_addr(&sBar[warp][s]);
			uint32_t done = 0;
			while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
			const int first = (int)(sOffset[qr[s]] + qc[s] * 128);
#pragma unroll
			for (int k = 0; k < 4; k++)
				asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
					:: "l"((uint64_t)&map), "r"(smem_addr(stage + s * 4096)), "r"(0), "r"(k), "r"(first) : "memory");
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			// refill this stage with the warp's next chunk once its stores have read it
			bool refilled = false;
			for (; issueJ < items && !refilled; issueJ += stride) {
				if (!valid(issueJ, r, c)) continue;
				asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
				qr[s] = r; qc[s] = c;
				issue_load(s, r, c);
				refilled = true;
			}
			if (!refilled) filled--;
			if (++s == kPeerStages) { s = 0; phase ^= 1u; }
		}
		asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
	}
	__syncwarp();
	// the tails: run r's last (count % 128) records, one run per warp at a time, through the lanes
	for (uint64_t r = gw; r < runs.numRanks; r += stride) {
		const unsigned long long count = sOffset[r + 1] - sOffset[r];
		const unsigned long long base = count / 128 * 128;
		const float4 *src = runs.buf[r] + 2;
		for (unsigned long long i = base + lane; i < count; i += 32) {
			float4 a, b;
			asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
				: "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(src + 2ull * i));
			float4 *d = out + 8ull * (sOffset[r] + i);
			st_record(d, a, b); st_record(d + 2, a, b); st_record(d + 4, a, b); st_record(d + 6, a, b);
		}
	}
}

// ---------------------------------------------------------------------------
// Launch wrappers
// ---------------------------------------------------------------------------

static inline uint32_t div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

void launch_fill_jobs(cudaStream_t st, uint32_t *dst, const FillJob *jobs, uint32_t numJobs)
{
	if (!numJobs) return;
	uint32_t blocks = div_up((uint64_t)numJobs * 32, 256);
	if (blocks > 4096) blocks = 4096;
	launch_chain(k_fill_jobs, blocks, 256, 0, st, dst, jobs, numJobs);
}

void launch_scatter_words(cudaStream_t st, void *dst, const uint32_t *idx, const void *src, uint32_t n, uint32_t words)
{
	if (!n) return;
	uint32_t blocks = div_up((uint64_t)n * words, 256);
	if (blocks > 8192) blocks = 8192;
	launch_chain(k_scatter_words, blocks, 256, 0, st, (uint32_t*)dst, idx, (const uint32_t*)src, n, words);
}

void launch_plan(cudaStream_t st, const TablePtrs &tab, const SystemWork *work, uint32_t numWork, const ActiveEntry *active, uint32_t stamp)
{
	if (!numWork) return;
	launch_chain(k_plan, div_up((uint64_t)numWork * 32, 128), 128, 0, st, tab, work, numWork, active, stamp);
}

void launch_emit_warp(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, const SystemWork *work, const ActiveEntry *active, uint32_t numActive, uint32_t round, uint32_t stamp)
{
	if (!numActive) return;
	launch_chain(k_emit_warp, div_up((uint64_t)numActive * 32, 256), 256, 0, st, pool, tab, work, active, numActive, round, stamp);
}

void launch_emit_burst(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, const ActiveEntry *active, const BurstJob *jobs, uint32_t numJobs, uint32_t totalBlocks)
{
	if (!numJobs || !totalBlocks) return;
	launch_chain(k_emit_burst, totalBlocks, 256, 0, st, pool, tab, active, jobs, numJobs);
}

void launch_integrate(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, uint32_t numPoolTiles, uint32_t stamp, uint32_t round,
	uint32_t epoch, bool sortMode, bool gravityPoints, bool tinyEffects)
{
	if (!numPoolTiles) return;
	(void)epoch;
#define SPR_LAUNCH_PASS(GP, TINY, KEYS) launch_chain(k_integrate_pass<GP, TINY, KEYS>, numPoolTiles, 256, 0, st, pool, tab, stamp, round)
	int variant = (gravityPoints ? 4 : 0) | (tinyEffects ? 2 : 0) | (sortMode ? 1 : 0);
	switch (variant) {
	case 0: SPR_LAUNCH_PASS(false, false, false); break;
	case 1: SPR_LAUNCH_PASS(false, false, true); break;
	case 2: SPR_LAUNCH_PASS(false, true, false); break;
	case 3: SPR_LAUNCH_PASS(false, true, true); break;
	case 4: SPR_LAUNCH_PASS(true, false, false); break;
	case 5: SPR_LAUNCH_PASS(true, false, true); break;
	case 6: SPR_LAUNCH_PASS(true, true, false); break;
	default: SPR_LAUNCH_PASS(true, true, true); break;
	}
#undef SPR_LAUNCH_PASS
}

void launch_dead_append(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, uint32_t numPoolTiles, uint32_t stamp, uint32_t round, uint32_t epoch, bool aliveScan,
	uint64_t &ticketBase)
{
	if (!numPoolTiles) return;
	const uint32_t blocks = div_up(numPoolTiles, 4);
	if (aliveScan) launch_chain(k_dead_append<true>, blocks, 128, 0, st, pool, tab, stamp, round, epoch, numPoolTiles, (unsigned long long)ticketBase);
	else launch_chain(k_dead_append<false>, blocks, 128, 0, st, pool, tab, stamp, round, epoch, numPoolTiles, (unsigned long long)ticketBase);
	ticketBase += (uint64_t)blocks * 4; // one ticket per launched warp
}

// Per-device kernel attributes (dynamic shared memory of the fused stream kernel) and its resident
// CTA count; called once per context, after cudaSetDevice.
void configure_stream_fused(int ctasPerSm[2])
{
	const size_t smem = 2 * 8 * 256 * sizeof(float4);
	cudaFuncSetAttribute(k_stream_fused<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	cudaFuncSetAttribute(k_stream_fused<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	cudaFuncSetAttribute(k_stream_fused<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	cudaFuncSetAttribute(k_stream_fused<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm[0], k_stream_fused<false, true>, 288, smem);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm[1], k_stream_fused<true, true>, 288, smem);
	if (ctasPerSm[0] < 1) ctasPerSm[0] = 1;
	if (ctasPerSm[1] < 1) ctasPerSm[1] = 1;
}

// Tensor map of the vertex pool for the bulk stores of k_stream_fused: [slots][4 copies][8 floats], box [128][1][8].
// cuTensorMapEncodeTiled is a driver entry point; it is fetched through the runtime so that the library links no libcuda.
bool encode_vertex_map(VertexTensorMap *out, void *vertices, uint64_t slots)
{
	typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
		CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	static_assert(sizeof(VertexTensorMap) == sizeof(CUtensorMap) && alignof(VertexTensorMap) >= alignof(CUtensorMap), "opaque tensor map");
	static EncodeFn fn = [] {
		void *p = nullptr; cudaDriverEntryPointQueryResult qr;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) p = nullptr;
		return (EncodeFn)p;
	}();
	if (!fn || !vertices || !slots || slots > 0x7fffffffull) return false; // tensor coordinates are signed 32-bit
	const cuuint64_t dims[3] = { 8, 4, slots };
	const cuuint64_t strides[2] = { 32, 128 };
	const cuuint32_t box[3] = { 8, 1, kGranuleSlots };
	const cuuint32_t elemStrides[3] = { 1, 1, 1 };
	return fn((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, vertices, dims, strides, box, elemStrides,
		CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

uint32_t launch_stream_fused(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, uint32_t numPoolTiles, uint32_t stamp, uint32_t round,
	uint32_t epoch, bool gravityPoints, uint32_t numSms, const int ctasPerSm[2], uint64_t &ticketBase, const VertexTensorMap *vertexMap)
{
	if (!numPoolTiles) return 0;
	const size_t smem = 2 * 8 * 256 * sizeof(float4);
	const int v = gravityPoints ? 1 : 0;
	uint32_t grid = numSms * (uint32_t)ctasPerSm[v]; // persistent: one wave of CTAs when the device is otherwise idle (tiles go by ticket, so fewer resident CTAs are fine too)
	if (grid > numPoolTiles) grid = numPoolTiles;
	CUtensorMap map;
	memset(&map, 0, sizeof(map));
	if (vertexMap) memcpy(&map, vertexMap, sizeof(map));
	const unsigned long long tb = (unsigned long long)ticketBase;
	if (vertexMap) {
		if (v) launch_chain(k_stream_fused<true, true>, grid, 288, smem, st, pool, tab, stamp, round, epoch, numPoolTiles, tb, map);
		else launch_chain(k_stream_fused<false, true>, grid, 288, smem, st, pool, tab, stamp, round, epoch, numPoolTiles, tb, map);
	} else {
		if (v) launch_chain(k_stream_fused<true, false>, grid, 288, smem, st, pool, tab, stamp, round, epoch, numPoolTiles, tb, map);
		else launch_chain(k_stream_fused<false, false>, grid, 288, smem, st, pool, tab, stamp, round, epoch, numPoolTiles, tb, map);
	}
	ticketBase += (uint64_t)numPoolTiles + grid; // every CTA claims until its ticket runs past the last tile
	return grid;
}

void launch_expand_stream(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, uint32_t numPoolTiles, uint32_t stamp)
{
	if (!numPoolTiles) return;
	launch_chain(k_expand_stream, numPoolTiles, 256, 0, st, pool, tab, stamp);
}

void launch_finalize(cudaStream_t st, const TablePtrs &tab, const ActiveEntry *active, uint32_t numActive, uint32_t stamp)
{
	if (!numActive) return;
	launch_chain(k_finalize, div_up((uint64_t)numActive * 16, 256), 256, 0, st, tab, active, numActive, stamp);
}

void launch_query_frustum(cudaStream_t st, const TablePtrs &tab, uint32_t numGlobals, uint32_t systemIdx, const float frustum[32],
	uint32_t *outIds, uint32_t capacity, uint32_t *outCount)
{
	if (!numGlobals) return;
	FrustumDev f;
	memcpy(&f, frustum, sizeof(f));
	k_query_frustum<<<div_up(numGlobals, 256), 256, 0, st>>>(tab, numGlobals, systemIdx, f, outIds, capacity, outCount);
}

void launch_gather_render(cudaStream_t st, const TablePtrs &tab, const uint32_t *effects, uint32_t n, RenderGather *out)
{
	if (!n) return;
	k_gather_render<<<div_up(n, 128), 128, 0, st>>>(tab, effects, n, out);
}

void launch_run_offsets(cudaStream_t st, const TablePtrs &tab, const uint32_t *effects, uint32_t n, uint64_t *offsets, uint32_t *counts)
{
	k_run_offsets<<<1, 1024, 0, st>>>(tab, effects, n, offsets, counts);
}

void launch_pack_runs(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, const uint32_t *effects, uint32_t n,
	const uint64_t *offsets, float4 *out, uint64_t capacityRecords, uint32_t numSms, bool fromState)
{
	if (!n) return;
	if (fromState) k_pack_runs<true><<<numSms * 8, 256, 0, st>>>(pool, tab, effects, n, offsets, out, capacityRecords);
	else k_pack_runs<false><<<numSms * 8, 256, 0, st>>>(pool, tab, effects, n, offsets, out, capacityRecords);
}

void launch_pack_runs_buf(cudaStream_t st, const PoolPtrs &pool, const TablePtrs &tab, const uint32_t *effects, uint32_t n,
	const uint64_t *offsets, void *buf, uint64_t capacityRecords, uint32_t numSms, bool fromState)
{
	if (fromState) k_pack_runs_buf<true><<<numSms * 8, 256, 0, st>>>(pool, tab, effects, n, offsets, (float4*)buf, capacityRecords);
	else k_pack_runs_buf<false><<<numSms * 8, 256, 0, st>>>(pool, tab, effects, n, offsets, (float4*)buf, capacityRecords);
}

void launch_expand_from_peers(cudaStream_t st, void *const *bufs, uint32_t numRanks, uint32_t myRank, float4 *out, uint64_t capacityRecords, uint64_t *totalOut, uint32_t numCtas, bool smallShare)
{
	PeerRuns runs;
	runs.numRanks = numRanks;
	runs.myRank = myRank;
	for (uint32_t r = 0; r < 16; r++) runs.buf[r] = r < numRanks ? (const float4*)bufs[r] : nullptr;
	// Bulk-async variant for a kernel that shares the SMs with the next frame's simulation (smallShare: at most two CTAs per SM,
	// sprContextSetAssemblyShare): 128-thread CTAs, one pipeline per warp, 6 K registers per CTA instead of 18 K and none of them
	// holding data in flight -- 8 GPUs, deferred expansion: 3.19 -> 3.05 ms per step at one CTA per SM. With the GPU to itself the
	// LSU kernel's 256 threads x 4 records pull faster (2.97 against 3.12 ms): it stays the kernel of the larger shares.
	// SPR_PEER_TMA=0 / 1 forces one or the other (A/B).
	static const int forced = [] { const char *e = getenv("SPR_PEER_TMA"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
	const bool tma = forced >= 0 ? forced == 1 : smallShare;
	VertexTensorMap vm;
	if (tma && capacityRecords >= 128 && encode_vertex_map(&vm, out, capacityRecords)) {
		CUtensorMap map;
		memcpy(&map, &vm, sizeof(map));
		static bool configured = false;
		const size_t smem = (size_t)kPeerWarps * kPeerStages * 4096;
		if (!configured) { cudaFuncSetAttribute(k_expand_from_peers_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = true; }
		k_expand_from_peers_tma<<<numCtas, kPeerWarps * 32, smem, st>>>(runs, out, capacityRecords, (unsigned long long*)totalOut, map);
		return;
	}
	k_expand_from_peers<<<numCtas, 256, 0, st>>>(runs, out, capacityRecords, (unsigned long long*)totalOut); // persistent: grid-stride over 1024-record chunks
}

void launch_expand_runs(cudaStream_t st, const float4 *records, uint64_t numRecords, float4 *out, uint32_t numSms)
{
	if (!numRecords) return;
	uint32_t blocks = div_up(numRecords * 8, 256);
	if (blocks > numSms * 16) blocks = numSms * 16;
	k_expand_runs<<<blocks, 256, 0, st>>>(records, numRecords, out);
}

} // namespace spr
