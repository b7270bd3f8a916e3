/*
 * spear_oracle.c -- TEST INFRASTRUCTURE ONLY (kind "port", see oracle_api.h).
 *
 * Plain-C, single-threaded, scalar restatement of the reference's particle hot
 * path. Every function cites the reference lines it follows (paths relative to
 * the reference tree). It exists so that parity can be checked on machines
 * where the reference itself is not available, and it is itself pinned against
 * the reference (oracle/_ref, built from the reference's own sources) by
 * tests/test_oracle.py and by the golden vectors under tests/golden/.
 *
 * Parity status: PINNED -- against outputs of the reference run in this
 * container (the reference ships no tests or golden vectors of its own,
 * SURVEY.md section 4); generator: tests/golden/make_golden.py.
 *
 * Arithmetic rules: no FMA contraction (build with
 * -ffp-contract=off), IEEE sqrt/div, SSE min/max operand semantics,
 * round-to-nearest-even in approxCosSin2 (x86 _mm_round_ps, src/sf/Float4.h:188).
 * sf::Float4::rsqrt follows the forced-scalar branch 1/sqrt (src/sf/Float4.h:420).
 *
 * The product (spear_b200/) never links, loads or calls this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "oracle_api.h"

#define HUGE_LIFE 1e20f      /* ParticleSystem.cpp:27 */
#define HUGE_LIFE_CMP 1e19f  /* ParticleSystem.cpp:28 */
#define MAX_CACHE_FRAMES 16  /* ParticleSystem.cpp:25 */
#define MAX_PARTICLES_PER_FRAME (4 * 1024) /* ParticleSystem.cpp:24 */

#define F_PI 3.14159265358979323846f   /* src/sf/Base.h:167 */
#define F_2PI 6.28318530717958647692f  /* src/sf/Base.h:168 */

/* -- sf::min / sf::max for float on x86: minss / maxss (src/sf/Base.h:106-109) */
static float sf_minf(float a, float b) { return a < b ? a : b; }
static float sf_maxf(float a, float b) { return a > b ? a : b; }
static float sf_clampf(float v, float lo, float hi) { return sf_minf(sf_maxf(v, lo), hi); } /* Base.h:118-120 */

/* -- sf::Random (src/sf/Random.h:7-18, src/sf/Random.cpp:8-22) ------------- */

typedef struct Rng { uint64_t state, inc; } Rng;

static Rng rng_make(uint64_t state, uint64_t inc) { Rng r; r.state = state; r.inc = inc | 1; return r; } /* Random.h:11 */

static uint32_t rng_u32(Rng *r) /* Random.cpp:8-15 */
{
	uint64_t old = r->state;
	r->state = old * 6364136223846793005ULL + r->inc;
	uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
	uint32_t rot = (uint32_t)(old >> 59u);
	return (xorshifted >> rot) | (xorshifted << (((uint32_t)-(int32_t)rot) & 31));
}

static float rng_float(Rng *r) /* Random.cpp:17-20 */
{
	return (float)((double)rng_u32(r) * 2.3283064365386963e-10);
}

/* -- approximations (ParticleSystem.cpp:43-64) ------------------------------ */

static float approx_acos(float a) /* :43-47 */
{
	float x = fabsf(a);
	float v = -1.280827681872808f + 1.280827681872808f * sqrtf(1.0f - x) - 0.28996864492208857f * x;
	return F_PI * 0.5f - copysignf(v, a);
}

static float cossin_lane(float v, float bias) /* one lane of approxCosSin2, :49-60 */
{
	const float rcp2pi = 1.0f / F_2PI;
	float t = v;
	t *= rcp2pi;
	t += bias;
	t -= rintf(t - 0.25f) + 0.25f;
	t *= (fabsf(t) - 0.5f) * 16.0f;
	t += t * (fabsf(t) - 1.0f) * 0.225f;
	return t;
}

static float approx_cbrt(float a) /* :62-64 */
{
	return 1.5142669123810535f * sqrtf(a) - 0.5142669123810535f * a;
}

/* -- RandomVec3Imp (ParticleSystem.cpp:66-146) ------------------------------ */

enum { RV_BOX = 1, RV_SPHERE = 2, RV_ROTATION = 4 };

typedef struct RandomVec3Imp {
	uint32_t flags;
	SprVec3 offset, boxExtent;
	float thetaBias, thetaScale, cosPhiBias, cosPhiScale, radiusCubeScale, radiusScale;
	SprVec3 sphereScale;
	SprQuat rotation;
} RandomVec3Imp;

static SprQuat euler_to_quat_xyz(float x, float y, float z) /* src/sf/Quaternion.cpp:18-33 */
{
	x *= 0.5f; y *= 0.5f; z *= 0.5f;
	float cx = cosf(x), sx = sinf(x);
	float cy = cosf(y), sy = sinf(y);
	float cz = cosf(z), sz = sinf(z);
	SprQuat q;
	q.x = -cx*sy*sz + cy*cz*sx;
	q.y = cx*cz*sy + cy*sx*sz;
	q.z = cx*cy*sz - cz*sx*sy;
	q.w = cx*cy*cz + sx*sy*sz;
	return q;
}

static SprVec3 quat_rotate(SprQuat q, SprVec3 v) /* src/sf/Quaternion.h:66-75 */
{
	float xy = q.x*v.y - q.y*v.x;
	float xz = q.x*v.z - q.z*v.x;
	float yz = q.y*v.z - q.z*v.y;
	SprVec3 r;
	r.x = 2.0f * (+ q.w*yz + q.y*xy + q.z*xz) + v.x;
	r.y = 2.0f * (- q.x*xy - q.w*xz + q.z*yz) + v.y;
	r.z = 2.0f * (- q.x*xz - q.y*yz + q.w*xy) + v.z;
	return r;
}

static void rv_init(RandomVec3Imp *r, const SprRandomVec3 *sv) /* :91-116 */
{
	const float d2r = F_PI / 180.0f;
	memset(r, 0, sizeof(*r));
	r->offset = sv->offset;
	if (sv->boxExtent.x != 0.0f || sv->boxExtent.y != 0.0f || sv->boxExtent.z != 0.0f) {
		r->flags |= RV_BOX;
		r->boxExtent = sv->boxExtent;
	}
	if (sv->sphere.maxRadius > 0.0f) {
		r->flags |= RV_SPHERE;
		r->thetaBias = sv->sphere.minTheta * d2r;
		r->thetaScale = (sv->sphere.maxTheta - sv->sphere.minTheta) * d2r;
		r->cosPhiBias = cosf(sv->sphere.maxPhi * d2r);
		r->cosPhiScale = cosf(sv->sphere.minPhi * d2r) - r->cosPhiBias;
		r->radiusCubeScale = approx_cbrt(1.0f - sv->sphere.minRadius / sv->sphere.maxRadius);
		r->radiusScale = sv->sphere.maxRadius;
		r->sphereScale = sv->sphere.scale;
	}
	if (sv->rotation.x != 0.0f || sv->rotation.y != 0.0f || sv->rotation.z != 0.0f) {
		r->flags |= RV_ROTATION;
		r->rotation = euler_to_quat_xyz(sv->rotation.x * d2r, sv->rotation.y * d2r, sv->rotation.z * d2r);
	}
}

static SprVec3 rv_sample(const RandomVec3Imp *r, Rng *rng) /* :118-145 */
{
	SprVec3 p = r->offset;
	if (r->flags & RV_BOX) {
		float x = rng_float(rng), y = rng_float(rng), z = rng_float(rng); /* nextVec3, Random.cpp:29-35 */
		p.x += (x - 0.5f) * r->boxExtent.x;
		p.y += (y - 0.5f) * r->boxExtent.y;
		p.z += (z - 0.5f) * r->boxExtent.z;
	}
	if (r->flags & RV_SPHERE) {
		float u = rng_float(rng);
		float v = rng_float(rng);
		float w = rng_float(rng);
		float theta = r->thetaBias + u * r->thetaScale;
		float phi = approx_acos(r->cosPhiBias + v * r->cosPhiScale);
		float cosTheta = cossin_lane(theta, 0.0f), sinTheta = cossin_lane(theta, -0.25f);
		float cosPhi = cossin_lane(phi, 0.0f), sinPhi = cossin_lane(phi, -0.25f);
		float radius = approx_cbrt(1.0f - w * r->radiusCubeScale) * r->radiusScale;
		p.x += (sinPhi * cosTheta) * radius * r->sphereScale.x;
		p.y += (cosPhi) * radius * r->sphereScale.y;
		p.z += (sinPhi * sinTheta) * radius * r->sphereScale.z;
	}
	if (r->flags & RV_ROTATION) p = quat_rotate(r->rotation, p);
	return p;
}

/* -- B-spline LUT (src/client/BSpline.cpp:7-88) ----------------------------- */

static float bspline_eval(float a, float b, float c, float d, float t) /* BSpline.cpp:7-35 */
{
	float nt = 1.0f - t;
	float t2 = t*t;
	float t3 = t2*t;
	float nt2 = nt*nt;
	float nt3 = nt2*nt;
	float p = a * (nt3 / 6.0f);
	p += b * ((3.0f*t3 - 6.0f*t2 + 4.0f) / 6.0f);
	p += c * ((-3.0f*t3 + 3.0f*t2 + 3.0f*t + 1.0f) / 6.0f);
	p += d * (t3 / 6.0f);
	return p;
}

static void discretize_bspline_y(float *y, uint32_t ny, const SprVec2 *pts, uint32_t npts, float xMin, float xMax) /* BSpline.cpp:37-88 */
{
	if (ny == 0) return;
	if (ny == 1) { y[0] = pts[0].y; return; }
	float curX = xMin;
	float xStep = (xMax - xMin) / (float)(ny - 1);
	uint32_t curPoint = 0;
	int numPoints = (int)npts;
	for (int i = -1; i < numPoints; i++) {
		SprVec2 a = pts[i - 1 >= 0 ? i - 1 : 0];
		SprVec2 b = pts[i >= 0 ? i : 0];
		SprVec2 c = pts[i + 1 < numPoints ? i + 1 : numPoints - 1];
		SprVec2 d = pts[i + 2 < numPoints ? i + 2 : numPoints - 1];
		float maxX = (b.x + d.x + 4.0f * c.x) * (1.0f / 6.0f);
		while (curX <= maxX) {
			const uint32_t steps = 512;
			const float stepSize = 1.0f / (float)(steps - 1);
			uint32_t first = 0, count = steps;
			while (count > 0) {
				uint32_t step = count >> 1;
				uint32_t it = first + step;
				float t = (float)it * stepSize;
				float x = bspline_eval(a.x, b.x, c.x, d.x, t);
				if (x < curX) { first = it; count -= step + 1; } else { count = step; }
			}
			float t = (float)first * stepSize;
			y[curPoint] = bspline_eval(a.y, b.y, c.y, d.y, t);
			if (++curPoint == ny) return;
			curX += xStep;
		}
	}
	for (; curPoint < ny; curPoint++) y[curPoint] = curPoint > 0 ? y[curPoint - 1] : 0.0f;
}

static float linear_to_srgb(float x) /* src/sp/Srgb.cpp:15-22 */
{
	x = sf_clampf(x, 0.0f, 1.0f);
	if (x <= 0.0031308f) return 12.92f * x;
	return 1.055f * powf(x, (1.0f / 2.4f)) - 0.055f;
}

static uint8_t float_to_unorm(float v) { return (uint8_t)(v * 255.9f); } /* Srgb.cpp:10-13 */

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* -- state (ParticleSystem.cpp:148-263) ------------------------------------- */

typedef struct EffectType {
	const SprParticleSystemComponent *comp; /* svComponent, the pointer is the key (:30-41) */
	double renderOrder;
	uint32_t tiebreaker, refCount;
	float updateRadius, timeStep;
	RandomVec3Imp emitPosition, emitVelocity;
	SprVertexTypeUbo typeUbo;
	uint8_t lut[SPR_SPLINE_LUT_BYTES];
} EffectType;

typedef struct Slot { float px, py, pz, vx, vy, vz, life, seed; } Slot; /* one lane of Particle4 (:170-176) */

typedef struct Effect {
	uint32_t typeId;
	SprVec3 origin, prevOrigin;
	SprBounds3 gpuBounds;
	float timeDelta, timeStep;
	SprVec3 gravity;
	double lastUpdateTime;
	float emitterToWorld[4][4], prevEmitterToWorld[4][4];
	SprMat34 particlesToWorld;
	int stopEmit, firstEmit, instantDelete;
	float spawnTimer, emitOnTimer;
	uint32_t uploadByteOffset;
	uint64_t uploadFrameIndex;
	Slot *slots; uint32_t numBlocks, capBlocks;        /* particles */
	SprGpuParticle *gpu; uint32_t numGpu; size_t capGpu; /* gpuParticles */
	uint32_t *freeIdx; uint32_t numFree, capFree;      /* freeIndices */
	uint64_t simSteps;
	int inUse;
} Effect;

typedef struct SortedEffect { double renderOrder; uint32_t tiebreaker; float depth; uint32_t effectId; } SortedEffect; /* :225-239 */

typedef struct System {
	Effect *effects; uint32_t numEffects, capEffects;
	uint32_t *freeEffectIds; uint32_t numFreeEffects, capFreeEffects;
	EffectType *types; uint32_t numTypes, capTypes;
	uint32_t *freeTypeIds; uint32_t numFreeTypes, capFreeTypes;
	Rng initRng, rng;
	uint64_t bufferFrameIndex;
	uint32_t appendOffset[MAX_CACHE_FRAMES];
	SortedEffect *sorted; uint32_t numSorted, capSorted;
} System;

static void mat34_identity(SprMat34 *m) /* src/sf/Matrix.h:103-107 */
{
	memset(m, 0, sizeof(*m));
	m->v[0] = 1.0f; m->v[4] = 1.0f; m->v[8] = 1.0f;
}

static void mat34_to_colmajor44(const SprMat34 *m, float *dst) /* Matrix.h:131-136 */
{
	memcpy(dst + 0, m->v + 0, 12); dst[3] = 0.0f;
	memcpy(dst + 4, m->v + 3, 12); dst[7] = 0.0f;
	memcpy(dst + 8, m->v + 6, 12); dst[11] = 0.0f;
	memcpy(dst + 12, m->v + 9, 12); dst[15] = 1.0f;
}

static SprMat34 mat_world(SprVec3 t, SprQuat q, float scale) /* src/sf/Matrix.cpp:899-920 */
{
	float s2 = 2.0f * scale;
	float xx = q.x*q.x, xy = q.x*q.y, xz = q.x*q.z, xw = q.x*q.w;
	float yy = q.y*q.y, yz = q.y*q.z, yw = q.y*q.w;
	float zz = q.z*q.z, zw = q.z*q.w;
	SprMat34 m;
	m.v[0] = s2 * (- yy - zz + 0.5f);
	m.v[1] = s2 * (+ xy + zw);
	m.v[2] = s2 * (- yw + xz);
	m.v[3] = s2 * (- zw + xy);
	m.v[4] = s2 * (- xx - zz + 0.5f);
	m.v[5] = s2 * (+ xw + yz);
	m.v[6] = s2 * (+ xz + yw);
	m.v[7] = s2 * (- xw + yz);
	m.v[8] = s2 * (- xx - yy + 0.5f);
	m.v[9] = t.x; m.v[10] = t.y; m.v[11] = t.z;
	return m;
}

/* -- initEffectType (ParticleSystem.cpp:298-450) ---------------------------- */

static void bake_spline_channel(uint8_t *lut, int channel, const SprBSpline2 *sp, int oneMinus) /* :329-364 */
{
	float y[SPR_SPLINE_SAMPLE_RATE];
	if (sp->numPoints >= 1) {
		discretize_bspline_y(y, SPR_SPLINE_SAMPLE_RATE, sp->points, sp->numPoints, 0.0f, 1.0f);
		if (oneMinus) for (int i = 0; i < SPR_SPLINE_SAMPLE_RATE; i++) y[i] = 1.0f - y[i];
	} else {
		for (int i = 0; i < SPR_SPLINE_SAMPLE_RATE; i++) y[i] = 1.0f;
	}
	for (int i = 0; i < SPR_SPLINE_SAMPLE_RATE; i++) {
		lut[i * 4 + channel] = (uint8_t)clampi((int)(y[i] * 255.0f), 0, 255);
	}
}

static void init_effect_type(System *s, uint32_t typeId, const SprParticleSystemComponent *c)
{
	EffectType *type = &s->types[typeId];
	const float d2r = F_PI / 180.0f;
	type->updateRadius = c->updateRadius;
	type->timeStep = c->timeStep;
	rv_init(&type->emitPosition, &c->emitPosition);
	rv_init(&type->emitVelocity, &c->emitVelocity);
	type->renderOrder = c->renderOrder;
	type->tiebreaker = typeId;

	memset(&type->typeUbo, 0, sizeof(type->typeUbo)); /* :312 */
	type->typeUbo.u_FrameCount.x = (float)c->frameCount[0];
	type->typeUbo.u_FrameCount.y = (float)c->frameCount[1];
	type->typeUbo.u_FrameRate = c->frameRate;
	type->typeUbo.u_RotationControl.x = c->rotation * d2r;
	type->typeUbo.u_RotationControl.y = c->rotationVariance * d2r;
	type->typeUbo.u_RotationControl.z = c->spin * d2r;
	type->typeUbo.u_RotationControl.w = c->spinVariance * d2r;
	if (c->randomStartFrame) type->typeUbo.u_StartFrame = 1.0f;

	uint32_t atlasX = typeId / 256; /* SplineAtlasHeight, :255,:323-324 */
	uint32_t atlasY = typeId % 256;

	uint8_t *lut = type->lut;
	memset(lut, 0, SPR_SPLINE_LUT_BYTES); /* reference leaves gradient byte 3 uninitialised (:326,:398-399) */
	bake_spline_channel(lut, 0, &c->scaleSpline, 0);
	bake_spline_channel(lut, 1, &c->alphaSpline, 0);
	bake_spline_channel(lut, 2, &c->additiveSpline, 1);
	bake_spline_channel(lut, 3, &c->erosionSpline, 0);

	uint8_t *grad = lut + SPR_SPLINE_SAMPLE_RATE * 4; /* :366-401 */
	SprGradientPoint defaultGradient; defaultGradient.t = 1.0f; defaultGradient.color = c->gradient.defaultColor;
	const SprGradientPoint *points = c->gradient.points;
	uint32_t numPoints = c->gradient.numPoints;
	if (numPoints == 0) { points = &defaultGradient; numPoints = 1; }
	float splineStep = 1.0f / (float)(SPR_SPLINE_SAMPLE_RATE - 1);
	uint32_t pointIx = 0;
	for (uint32_t i = 0; i < SPR_SPLINE_SAMPLE_RATE; i++) {
		float t = (float)i * splineStep;
		while (pointIx < numPoints && t >= points[pointIx].t) pointIx++;
		SprVec3 col;
		if (pointIx == 0) col = points[0].color;
		else if (pointIx == numPoints) col = points[numPoints - 1].color;
		else {
			const SprGradientPoint *p0 = &points[pointIx - 1], *p1 = &points[pointIx];
			float relT = (t - p0->t) / (p1->t - p0->t);
			/* sf::lerp(Vec3): a * (1 - t) + b * t, src/sf/Vector.h:108 */
			col.x = p0->color.x * (1.0f - relT) + p1->color.x * relT;
			col.y = p0->color.y * (1.0f - relT) + p1->color.y * relT;
			col.z = p0->color.z * (1.0f - relT) + p1->color.z * relT;
		}
		uint8_t *dst = grad + i * 4;
		dst[0] = float_to_unorm(linear_to_srgb(col.x)); /* Srgb.cpp:48-53 */
		dst[1] = float_to_unorm(linear_to_srgb(col.y));
		dst[2] = float_to_unorm(linear_to_srgb(col.z));
	}

	uint32_t x = atlasX * SPR_SPLINE_SAMPLE_RATE; /* :403-413 */
	uint32_t y = atlasY * 2;
	float resX = (float)(SPR_SPLINE_SAMPLE_RATE * 8), resY = (float)(256 * 2);
	type->typeUbo.u_SplineMad.x = (float)(SPR_SPLINE_SAMPLE_RATE - 1) / resX;
	type->typeUbo.u_SplineMad.y = 1.0f / resY;
	type->typeUbo.u_SplineMad.z = ((float)x + 0.5f) / resX;
	type->typeUbo.u_SplineMad.w = ((float)y + 0.5f) / resY;
	type->typeUbo.u_ScaleBaseVariance.x = c->size;
	type->typeUbo.u_ScaleBaseVariance.y = c->sizeVariance;
	type->typeUbo.u_LifeTimeBaseVariance.x = c->lifeTime;
	type->typeUbo.u_LifeTimeBaseVariance.y = c->lifeTimeVariance * (1.0f / 16777216.0f);
}

/* -- simulateParticlesImp (ParticleSystem.cpp:452-680) ---------------------- */

static void free_push(Effect *e, uint32_t v)
{
	if (e->numFree == e->capFree) {
		e->capFree = e->capFree ? e->capFree * 2 : 16;
		e->freeIdx = (uint32_t*)realloc(e->freeIdx, sizeof(uint32_t) * e->capFree);
	}
	e->freeIdx[e->numFree++] = v;
}

static void simulate(System *s, Rng *rng, Effect *e, float dt)
{
	EffectType *type = &s->types[e->typeId];
	const SprParticleSystemComponent *comp = type->comp;

	uint32_t burstAmount = 0;
	if (e->firstEmit) { /* :457-464 */
		memcpy(e->prevEmitterToWorld, e->emitterToWorld, sizeof(e->emitterToWorld));
		burstAmount = comp->burstAmount;
		e->firstEmit = 0;
	}
	if (e->emitOnTimer >= 0.0f) { /* :466-471 */
		e->emitOnTimer -= dt;
		if (e->emitOnTimer < 0.0f) e->stopEmit = 1;
	}
	float drag = comp->drag;

	/* spawn loop, :476-524 */
	e->spawnTimer -= dt;
	int spawnsLeft = 20;
	while ((e->spawnTimer <= 0.0f && !e->stopEmit) || burstAmount > 0) {
		if (burstAmount > 0) {
			burstAmount--;
		} else {
			if (spawnsLeft-- <= 0) break;
			e->spawnTimer += comp->spawnTime + comp->spawnTimeVariance * rng_float(rng);
		}
		if (e->numFree == 0) { /* :487-494 */
			uint32_t base = e->numBlocks * 4;
			if (e->numBlocks == e->capBlocks) {
				uint32_t nc = e->capBlocks ? e->capBlocks * 2 : 16;
				e->slots = (Slot*)realloc(e->slots, sizeof(Slot) * 4 * nc);
				memset(e->slots + (size_t)e->capBlocks * 4, 0, sizeof(Slot) * 4 * (nc - e->capBlocks));
				e->capBlocks = nc;
			}
			for (uint32_t i = 0; i < 4; i++) e->slots[base + i].life = HUGE_LIFE;
			e->numBlocks++;
			for (uint32_t i = 0; i < 4; i++) free_push(e, base + i);
		}
		uint32_t index = e->freeIdx[--e->numFree];

		SprVec3 pos = rv_sample(&type->emitPosition, rng);
		SprVec3 vel = rv_sample(&type->emitVelocity, rng);

		if (comp->emitVelocityAttractorStrength != 0.0f) { /* :501-503, normalizeOrZero src/sf/Vector.h:101-105 */
			SprVec3 d;
			d.x = pos.x - comp->emitVelocityAttractorOffset.x;
			d.y = pos.y - comp->emitVelocityAttractorOffset.y;
			d.z = pos.z - comp->emitVelocityAttractorOffset.z;
			float len = sqrtf(d.x*d.x + d.y*d.y + d.z*d.z);
			if ((double)len > 1e-9) {
				float rcp = 1.0f / len;
				d.x *= rcp; d.y *= rcp; d.z *= rcp;
			} else {
				d.x = d.y = d.z = 0.0f;
			}
			vel.x += d.x * comp->emitVelocityAttractorStrength;
			vel.y += d.y * comp->emitVelocityAttractorStrength;
			vel.z += d.z * comp->emitVelocityAttractorStrength;
		}

		float emitT = sf_clampf(-e->spawnTimer / dt, 0.0f, 1.0f); /* :505 */
		float world[3];
		for (int k = 0; k < 3; k++) { /* :506-509, lanes x,y,z */
			float p = e->emitterToWorld[3][k] + (e->prevEmitterToWorld[3][k] - e->emitterToWorld[3][k]) * emitT;
			p += (e->emitterToWorld[0][k] + (e->prevEmitterToWorld[0][k] - e->emitterToWorld[0][k]) * emitT) * pos.x;
			p += (e->emitterToWorld[1][k] + (e->prevEmitterToWorld[1][k] - e->emitterToWorld[1][k]) * emitT) * pos.y;
			p += (e->emitterToWorld[2][k] + (e->prevEmitterToWorld[2][k] - e->emitterToWorld[2][k]) * emitT) * pos.z;
			world[k] = p;
		}
		Slot *p = &e->slots[index]; /* :513-522 */
		p->px = world[0]; p->py = world[1]; p->pz = world[2];
		p->vx = vel.x; p->vy = vel.y; p->vz = vel.z;
		p->life = 1.0f;
		p->seed = rng_float(rng) * 16777216.0f;
	}
	e->spawnTimer = sf_maxf(e->spawnTimer, 0.0f); /* :524 */

	/* integration, :526-666 */
	size_t needGpu = (size_t)e->numBlocks * 16;
	if (needGpu > e->capGpu) {
		size_t nc = e->capGpu ? e->capGpu : 64;
		while (nc < needGpu) nc *= 2;
		e->gpu = (SprGpuParticle*)realloc(e->gpu, sizeof(SprGpuParticle) * nc);
		e->capGpu = nc;
	}
	SprGpuParticle *dst = e->gpu;
	float pMin[4] = { HUGE_VALF, HUGE_VALF, HUGE_VALF, HUGE_VALF };
	float pMax[4] = { -HUGE_VALF, -HUGE_VALF, -HUGE_VALF, -HUGE_VALF };

	float lifeTime = comp->lifeTime;
	float lifeTimeVariance = comp->lifeTimeVariance * (1.0f / 16777216.0f);

	SprGravityPoint gp[SPR_MAX_GRAVITY_POINTS]; /* :539-550 */
	uint32_t numGp = comp->numGravityPoints < SPR_MAX_GRAVITY_POINTS ? comp->numGravityPoints : SPR_MAX_GRAVITY_POINTS;
	for (uint32_t i = 0; i < numGp; i++) {
		const SprGravityPoint *p = &comp->gravityPoints[i];
		float w[3];
		for (int k = 0; k < 3; k++) {
			float v = e->emitterToWorld[3][k] + (e->prevEmitterToWorld[3][k] - e->emitterToWorld[3][k]);
			v += e->emitterToWorld[0][k] * p->position.x;
			v += e->emitterToWorld[1][k] * p->position.y;
			v += e->emitterToWorld[2][k] * p->position.z;
			w[k] = v;
		}
		gp[i].position.x = w[0]; gp[i].position.y = w[1]; gp[i].position.z = w[2];
		gp[i].radius = sf_maxf(p->radius, 0.05f);
		gp[i].strength = -p->strength;
	}

	uint32_t numSlots = e->numBlocks * 4;
	for (uint32_t i = 0; i < numSlots; i++) {
		Slot *p = &e->slots[i];
		float life = p->life;
		if (life <= 0.0f) { /* :556-567 (per lane; NaN is neither) */
			free_push(e, i);
			life = HUGE_LIFE;
		}
		float seed = p->seed;
		float px = p->px, py = p->py, pz = p->pz;
		float vx = p->vx, vy = p->vy, vz = p->vz;

		float ax = e->gravity.x, ay = e->gravity.y, az = e->gravity.z; /* :574-579 */
		ax -= vx * drag;
		ay -= vy * drag;
		az -= vz * drag;

		for (uint32_t g = 0; g < numGp; g++) { /* :581-590 */
			float dx = px - gp[g].position.x;
			float dy = py - gp[g].position.y;
			float dz = pz - gp[g].position.z;
			float lenSq = (dx*dx + dy*dy + dz*dz) + gp[g].radius;
			float weight = (1.0f / sqrtf(lenSq)) / lenSq * gp[g].strength;
			ax += dx * weight;
			ay += dy * weight;
			az += dz * weight;
		}

		vx += ax * dt; vy += ay * dt; vz += az * dt; /* :592-596 */
		p->vx = vx; p->vy = vy; p->vz = vz;
		px += vx * dt; py += vy * dt; pz += vz * dt; /* :598-602 */
		p->px = px; p->py = py; p->pz = pz;
		life -= dt / (seed * lifeTimeVariance + lifeTime); /* :604 */
		p->life = life;

		if (life < HUGE_LIFE_CMP) { /* :611-665 */
			SprGpuParticle g;
			g.position.x = px; g.position.y = py; g.position.z = pz; g.life = life;
			g.velocity.x = vx; g.velocity.y = vy; g.velocity.z = vz; g.seed = seed;
			dst[0] = g; dst[1] = g; dst[2] = g; dst[3] = g;
			float v4[4] = { px, py, pz, life };
			for (int k = 0; k < 4; k++) {
				pMin[k] = pMin[k] < v4[k] ? pMin[k] : v4[k]; /* _mm_min_ps(pMin, v) */
				pMax[k] = pMax[k] > v4[k] ? pMax[k] : v4[k]; /* _mm_max_ps(pMax, v) */
			}
			dst += 4;
		}
	}

	memcpy(e->prevEmitterToWorld, e->emitterToWorld, sizeof(e->emitterToWorld)); /* :668 */
	e->numGpu = (uint32_t)(dst - e->gpu); /* :670 */

	e->gpuBounds.origin.x = (pMin[0] + pMax[0]) * 0.5f; /* :672-673 */
	e->gpuBounds.origin.y = (pMin[1] + pMax[1]) * 0.5f;
	e->gpuBounds.origin.z = (pMin[2] + pMax[2]) * 0.5f;
	e->gpuBounds.extent.x = (pMax[0] - pMin[0]) * 0.5f + comp->cullPadding;
	e->gpuBounds.extent.y = (pMax[1] - pMin[1]) * 0.5f + comp->cullPadding;
	e->gpuBounds.extent.z = (pMax[2] - pMin[2]) * 0.5f + comp->cullPadding;

	if (comp->localSpace) { /* :675-677, sf::transformBounds src/sf/Geometry.cpp:13-19, Matrix.cpp:579-586,:625-632 */
		const float *m = e->particlesToWorld.v;
		SprVec3 o = e->gpuBounds.origin, x = e->gpuBounds.extent, ro, rx;
		ro.x = m[0]*o.x + m[3]*o.y + m[6]*o.z + m[9];
		ro.y = m[1]*o.x + m[4]*o.y + m[7]*o.z + m[10];
		ro.z = m[2]*o.x + m[5]*o.y + m[8]*o.z + m[11];
		rx.x = fabsf(m[0])*x.x + fabsf(m[3])*x.y + fabsf(m[6])*x.z;
		rx.y = fabsf(m[1])*x.x + fabsf(m[4])*x.y + fabsf(m[7])*x.z;
		rx.z = fabsf(m[2])*x.x + fabsf(m[5])*x.y + fabsf(m[8])*x.z;
		e->gpuBounds.origin = ro; e->gpuBounds.extent = rx;
	}
	e->uploadFrameIndex = 0; /* :679 */
	e->simSteps++;
}

/* -- API (ParticleSystem.cpp:682-903) --------------------------------------- */

static void effect_reset(Effect *e) /* sf::reset(effect) -> default member initialisers, :186-218 */
{
	free(e->slots); free(e->gpu); free(e->freeIdx);
	memset(e, 0, sizeof(*e));
	e->firstEmit = 1;
	e->emitOnTimer = -1.0f;
	e->uploadFrameIndex = MAX_CACHE_FRAMES;
	mat34_identity(&e->particlesToWorld);
}

const char *orc_kind(void) { return "port"; }

void *orc_create(const SprSystemsDesc *desc) /* :265-267, :250 */
{
	System *s = (System*)calloc(1, sizeof(System));
	s->initRng = rng_make(1, 1);
	s->rng = rng_make(desc->seed[1], 581271);
	s->bufferFrameIndex = MAX_CACHE_FRAMES;
	return s;
}

void orc_destroy(void *sys)
{
	System *s = (System*)sys;
	for (uint32_t i = 0; i < s->numEffects; i++) { free(s->effects[i].slots); free(s->effects[i].gpu); free(s->effects[i].freeIdx); }
	free(s->effects); free(s->freeEffectIds); free(s->types); free(s->freeTypeIds); free(s->sorted);
	free(s);
}

static void release_type_imp(System *s, uint32_t typeId) /* :682-692 */
{
	EffectType *type = &s->types[typeId];
	if (--type->refCount == 0) {
		memset(type, 0, sizeof(*type));
		if (s->numFreeTypes == s->capFreeTypes) {
			s->capFreeTypes = s->capFreeTypes ? s->capFreeTypes * 2 : 16;
			s->freeTypeIds = (uint32_t*)realloc(s->freeTypeIds, sizeof(uint32_t) * s->capFreeTypes);
		}
		s->freeTypeIds[s->numFreeTypes++] = typeId;
	}
}

uint32_t orc_reserve_type(void *sys, const SprParticleSystemComponent *c) /* :696-720 */
{
	System *s = (System*)sys;
	uint32_t typeId = ~0u;
	for (uint32_t i = 0; i < s->numTypes; i++) if (s->types[i].comp == c) { typeId = i; break; }
	if (typeId == ~0u) {
		if (s->numFreeTypes > 0) {
			typeId = s->freeTypeIds[--s->numFreeTypes];
		} else {
			if (s->numTypes == s->capTypes) {
				s->capTypes = s->capTypes ? s->capTypes * 2 : 16;
				s->types = (EffectType*)realloc(s->types, sizeof(EffectType) * s->capTypes);
			}
			typeId = s->numTypes++;
			memset(&s->types[typeId], 0, sizeof(EffectType));
		}
		s->types[typeId].comp = c;
		init_effect_type(s, typeId, c);
	}
	s->types[typeId].refCount++;
	return typeId;
}

void orc_release_type(void *sys, const SprParticleSystemComponent *c) /* :722-728 */
{
	System *s = (System*)sys;
	for (uint32_t i = 0; i < s->numTypes; i++) if (s->types[i].comp == c) { release_type_imp(s, i); return; }
}

uint32_t orc_add_effect(void *sys, uint32_t entityId, uint8_t componentIndex,
	const SprParticleSystemComponent *c, const SprTransform *transform) /* :730-766 */
{
	System *s = (System*)sys;
	(void)entityId; (void)componentIndex;
	uint32_t effectId;
	if (s->numFreeEffects > 0) {
		effectId = s->freeEffectIds[--s->numFreeEffects];
	} else {
		if (s->numEffects == s->capEffects) {
			s->capEffects = s->capEffects ? s->capEffects * 2 : 16;
			s->effects = (Effect*)realloc(s->effects, sizeof(Effect) * s->capEffects);
		}
		effectId = s->numEffects++;
		memset(&s->effects[effectId], 0, sizeof(Effect));
		effect_reset(&s->effects[effectId]);
	}
	uint32_t typeId = orc_reserve_type(sys, c);
	EffectType *type = &s->types[typeId];
	Effect *e = &s->effects[effectId];
	e->inUse = 1;
	e->typeId = typeId;
	e->timeStep = type->timeStep * (1.0f + rng_float(&s->initRng) * 0.1f); /* :744 */
	e->gravity = c->gravity;
	e->prevOrigin = e->origin = transform->position;
	e->emitOnTimer = c->emitterOnTime;
	e->instantDelete = c->instantDelete;
	SprMat34 entityToWorld = mat_world(transform->position, transform->rotation, transform->scale);
	if (c->localSpace) { /* :755-760 */
		SprMat34 id; mat34_identity(&id);
		mat34_to_colmajor44(&id, &e->emitterToWorld[0][0]);
		e->particlesToWorld = entityToWorld;
	} else {
		mat34_to_colmajor44(&entityToWorld, &e->emitterToWorld[0][0]);
	}
	return effectId;
}

void orc_update_transform(void *sys, uint32_t effectId, const SprTransformUpdate *update) /* :768-784 */
{
	System *s = (System*)sys;
	Effect *e = &s->effects[effectId];
	EffectType *type = &s->types[e->typeId];
	if (type->comp->localSpace) e->particlesToWorld = update->entityToWorld;
	else mat34_to_colmajor44(&update->entityToWorld, &e->emitterToWorld[0][0]);
	e->origin = update->transform.position;
}

int orc_prepare_for_remove(void *sys, uint32_t effectId, const SprFrameArgs *args) /* :786-794 */
{
	System *s = (System*)sys;
	Effect *e = &s->effects[effectId];
	e->stopEmit = 1;
	return e->numGpu == 0 || args->gameTime - e->lastUpdateTime > 2.0 || e->instantDelete;
}

void orc_remove(void *sys, uint32_t effectId) /* :796-806 */
{
	System *s = (System*)sys;
	Effect *e = &s->effects[effectId];
	release_type_imp(s, e->typeId);
	effect_reset(e);
	if (s->numFreeEffects == s->capFreeEffects) {
		s->capFreeEffects = s->capFreeEffects ? s->capFreeEffects * 2 : 16;
		s->freeEffectIds = (uint32_t*)realloc(s->freeEffectIds, sizeof(uint32_t) * s->capFreeEffects);
	}
	s->freeEffectIds[s->numFreeEffects++] = effectId;
}

void orc_update(void *sys, const uint32_t *activeIds, uint32_t numActive, const SprFrameArgs *args) /* :808-822 */
{
	System *s = (System*)sys;
	float clampedDt = sf_minf(args->dt, 0.1f);
	for (uint32_t i = 0; i < numActive; i++) {
		Effect *e = &s->effects[activeIds[i]];
		e->lastUpdateTime = args->gameTime;
		e->timeDelta += clampedDt;
		while (e->timeDelta >= e->timeStep) {
			simulate(s, &s->rng, e, e->timeStep);
			e->timeDelta -= e->timeStep;
		}
	}
}

static int frustum_intersects(const SprFrustum *f, const SprBounds3 *b) /* src/sf/Frustum.h:60-75 */
{
	float o[3] = { b->origin.x, b->origin.y, b->origin.z };
	float e[3] = { b->extent.x, b->extent.y, b->extent.z };
	for (int lane = 0; lane < 4; lane++) {
		float so = f->side[3][lane], co = f->caps[3][lane];
		for (int k = 0; k < 3; k++) {
			float sv = f->side[k][lane], cv = f->caps[k][lane];
			so += sv * o[k]; co += cv * o[k];
			so += fabsf(sv) * e[k]; co += fabsf(cv) * e[k];
		}
		float m = so < co ? so : co; /* _mm_min_ps(so, co) */
		if (!(m > 0.0f)) return 0;
	}
	return 1;
}

static int sorted_cmp(const void *pa, const void *pb) /* SortedEffect::operator<, :232-238 */
{
	const SortedEffect *a = (const SortedEffect*)pa, *b = (const SortedEffect*)pb;
	if (a->renderOrder != b->renderOrder) return a->renderOrder < b->renderOrder ? -1 : 1;
	if (a->tiebreaker != b->tiebreaker) return a->tiebreaker < b->tiebreaker ? -1 : 1;
	if (a->depth != b->depth) return a->depth < b->depth ? -1 : 1;
	if (a->effectId != b->effectId) return a->effectId < b->effectId ? -1 : 1;
	return 0;
}

uint32_t orc_render(void *sys, const uint32_t *visibleIds, uint32_t numVisible,
	const SprRenderArgs *args, SprDrawRecord *out, uint32_t capacity, uint32_t *numSorted) /* :824-903 */
{
	System *s = (System*)sys;
	uint32_t frameParticlesLeft = MAX_PARTICLES_PER_FRAME;
	s->bufferFrameIndex++;
	uint64_t frame = s->bufferFrameIndex;

	if (s->capSorted < numVisible) {
		s->capSorted = numVisible;
		s->sorted = (SortedEffect*)realloc(s->sorted, sizeof(SortedEffect) * (numVisible ? numVisible : 1));
	}
	s->numSorted = 0;
	for (uint32_t i = 0; i < numVisible; i++) { /* :835-846 */
		uint32_t effectId = visibleIds[i];
		Effect *e = &s->effects[effectId];
		EffectType *type = &s->types[e->typeId];
		uint32_t numParticles = e->numGpu / 4;
		if (numParticles == 0) continue;
		if (!frustum_intersects(&args->frustum, &e->gpuBounds)) continue;
		SortedEffect *se = &s->sorted[s->numSorted++];
		se->renderOrder = type->renderOrder;
		se->tiebreaker = type->tiebreaker;
		float dx = args->cameraPosition.x - e->origin.x, dy = args->cameraPosition.y - e->origin.y, dz = args->cameraPosition.z - e->origin.z;
		se->depth = dx*dx + dy*dy + dz*dz;
		se->effectId = effectId;
	}
	qsort(s->sorted, s->numSorted, sizeof(SortedEffect), sorted_cmp); /* :848; strict total order */
	if (numSorted) *numSorted = s->numSorted;

	uint32_t n = 0;
	uint32_t prevTypeId = ~0u;
	for (uint32_t i = 0; i < s->numSorted; i++) { /* :851-902 */
		uint32_t effectId = s->sorted[i].effectId;
		Effect *e = &s->effects[effectId];
		uint32_t numParticles = e->numGpu / 4;
		if (frame - e->uploadFrameIndex >= MAX_CACHE_FRAMES) {
			if (numParticles >= frameParticlesLeft) break; /* `return`, :862 */
			frameParticlesLeft -= numParticles;
			uint32_t *off = &s->appendOffset[(uint32_t)frame % MAX_CACHE_FRAMES];
			e->uploadByteOffset = *off;
			*off += e->numGpu * (uint32_t)sizeof(SprGpuParticle);
			e->uploadFrameIndex = frame;
		}
		if (n < capacity) {
			SprDrawRecord *r = &out[n];
			memset(r, 0, sizeof(*r));
			r->effectId = effectId;
			r->typeId = e->typeId;
			r->numParticles = numParticles;
			r->vertexByteOffset = e->uploadByteOffset;
			mat34_to_colmajor44(&e->particlesToWorld, r->instanceUbo.u_ParticleToWorld); /* :871-876 */
			memcpy(r->instanceUbo.u_WorldToClip, args->worldToClip.v, sizeof(float) * 16);
			r->instanceUbo.u_Aspect = args->viewToClip.v[0] / args->viewToClip.v[5];
			r->instanceUbo.u_InvDelta = e->timeStep - e->timeDelta;
			r->instanceUbo.u_VisFogMad = args->visFogWorldMad;
			if (e->typeId != prevTypeId) { r->typeUboChanged = 1; prevTypeId = e->typeId; } /* :880-883 */
			n++;
		}
	}
	return n;
}

void orc_effect_info(void *sys, uint32_t effectId, SprEffectInfo *out)
{
	System *s = (System*)sys;
	const Effect *e = &s->effects[effectId];
	memset(out, 0, sizeof(*out));
	out->typeId = e->typeId;
	out->slotCount = e->numBlocks * 4;
	out->aliveCount = e->numGpu / 4;
	out->freeCount = e->numFree;
	out->timeStep = e->timeStep;
	out->timeDelta = e->timeDelta;
	out->spawnTimer = e->spawnTimer;
	out->emitOnTimer = e->emitOnTimer;
	out->stopEmit = (uint8_t)e->stopEmit;
	out->firstEmit = (uint8_t)e->firstEmit;
	out->instantDelete = (uint8_t)e->instantDelete;
	out->gpuBounds = e->gpuBounds;
	out->lastUpdateTime = e->lastUpdateTime;
	out->simSteps = e->simSteps;
}

uint32_t orc_read_free(void *sys, uint32_t effectId, uint32_t *out, uint32_t capacity)
{
	System *s = (System*)sys;
	const Effect *e = &s->effects[effectId];
	uint32_t n = e->numFree < capacity ? e->numFree : capacity;
	if (n) memcpy(out, e->freeIdx, sizeof(uint32_t) * n);
	return e->numFree;
}

uint32_t orc_read_stream(void *sys, uint32_t effectId, SprGpuParticle *out, uint32_t capacityVertices)
{
	System *s = (System*)sys;
	const Effect *e = &s->effects[effectId];
	uint32_t n = e->numGpu < capacityVertices ? e->numGpu : capacityVertices;
	if (n) memcpy(out, e->gpu, sizeof(SprGpuParticle) * (size_t)n);
	return e->numGpu;
}

uint32_t orc_read_slots(void *sys, uint32_t effectId, SprGpuParticle *out, uint32_t capacitySlots)
{
	System *s = (System*)sys;
	const Effect *e = &s->effects[effectId];
	uint32_t total = e->numBlocks * 4;
	uint32_t n = total < capacitySlots ? total : capacitySlots;
	for (uint32_t i = 0; i < n; i++) {
		const Slot *p = &e->slots[i];
		out[i].position.x = p->px; out[i].position.y = p->py; out[i].position.z = p->pz; out[i].life = p->life;
		out[i].velocity.x = p->vx; out[i].velocity.y = p->vy; out[i].velocity.z = p->vz; out[i].seed = p->seed;
	}
	return total;
}

void orc_rng_state(void *sys, uint64_t outStateInc[2])
{
	System *s = (System*)sys;
	outStateInc[0] = s->rng.state;
	outStateInc[1] = s->rng.inc;
}

void orc_type_lut(void *sys, uint32_t typeId, uint8_t out[SPR_SPLINE_LUT_BYTES])
{
	System *s = (System*)sys;
	memcpy(out, s->types[typeId].lut, SPR_SPLINE_LUT_BYTES);
}

void orc_type_ubo(void *sys, uint32_t typeId, SprVertexTypeUbo *out)
{
	System *s = (System*)sys;
	*out = s->types[typeId].typeUbo;
}

void orc_make_frustum(const SprMat44 *m, float clipNearW, SprFrustum *out) /* src/sf/Frustum.h:16-40, scalar rsqrt */
{
	/* rows of the column-major matrix: r_k[lane j] = cols[j].v[k] after transpose4 */
	float r[4][4];
	for (int k = 0; k < 4; k++) for (int j = 0; j < 4; j++) r[k][j] = m->v[j * 4 + k];
	{
		float p[4][4]; /* planes a,b,c,d before transpose: plane index, component lane */
		for (int j = 0; j < 4; j++) {
			p[0][j] = r[3][j] - r[0][j];
			p[1][j] = r[3][j] + r[0][j];
			p[2][j] = r[3][j] - r[1][j];
			p[3][j] = r[3][j] + r[1][j];
		}
		/* after transpose4: a = x components of the 4 planes, ... */
		for (int pl = 0; pl < 4; pl++) {
			float rcpLen = 1.0f / sqrtf(p[pl][0]*p[pl][0] + p[pl][1]*p[pl][1] + p[pl][2]*p[pl][2]);
			for (int comp = 0; comp < 4; comp++) out->side[comp][pl] = p[pl][comp] * rcpLen;
		}
	}
	{
		float p[4][4];
		for (int j = 0; j < 4; j++) {
			p[0][j] = r[3][j] * -clipNearW + r[2][j];
			p[1][j] = r[3][j] - r[2][j];
			p[2][j] = p[0][j];
			p[3][j] = p[1][j];
		}
		for (int pl = 0; pl < 4; pl++) {
			float rcpLen = 1.0f / sqrtf(p[pl][0]*p[pl][0] + p[pl][1]*p[pl][1] + p[pl][2]*p[pl][2]);
			for (int comp = 0; comp < 4; comp++) out->caps[comp][pl] = p[pl][comp] * rcpLen;
		}
	}
}

double orc_bench_steps(void *sys, const uint32_t *activeIds, uint32_t numActive,
	const SprFrameArgs *args, uint32_t steps, uint64_t *particleSteps)
{
	System *s = (System*)sys;
	SprFrameArgs fa = *args;
	uint64_t total = 0;
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (uint32_t k = 0; k < steps; k++) {
		fa.frameIndex++;
		fa.gameTime += (double)fa.dt;
		orc_update(sys, activeIds, numActive, &fa);
		for (uint32_t i = 0; i < numActive; i++) total += s->effects[activeIds[i]].numGpu / 4;
	}
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (particleSteps) *particleSteps += total;
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
