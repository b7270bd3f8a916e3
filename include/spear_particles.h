/*
 * spear_particles.h -- C ABI of the B200-native replacement for Spear's CPU
 * particle hot path (cl::ParticleSystem, src/client/ParticleSystem.{h,cpp}).
 *
 * Everything here is plain C: POD structs, raw pointers and sizes. No torch,
 * no C++ types. The same POD descriptors are consumed by the product library
 * (spear_b200/csrc -> libspear_particles.so, CUDA sm_100a) and by the CPU
 * oracles under oracle/ (test infrastructure only).
 *
 * Every entry point cites the reference interface it replaces (paths relative
 * to the reference tree, file:line).
 *
 * Conventions kept from the reference:
 *  - single host thread drives one SprContext; update then render alternate;
 *  - effect types are keyed by the descriptor POINTER (ParticleSystem.cpp:30-41,
 *    :699-700): the SprParticleSystemComponent passed to sprReserveEffectType /
 *    sprAddEffect must stay alive and must not move while the type is reserved;
 *  - effect / type ids are dense u32 recycled LIFO (:242-245, :703-707, :733-737);
 *  - the reference has no error channel; every call here returns SPR_OK (0) on
 *    the reference-compatible subset and a negative SprStatus otherwise
 *    (bad argument, CUDA failure). There is NO CPU fallback: if the CUDA
 *    runtime or a B200-class device is missing, sprContextCreate fails.
 */
#ifndef SPEAR_PARTICLES_H
#define SPEAR_PARTICLES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SPR_API __declspec(dllexport)
#else
#define SPR_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------------ */
/* Status                                                                    */
/* ------------------------------------------------------------------------ */

typedef enum SprStatus {
	SPR_OK = 0,
	SPR_ERR_INVALID_ARGUMENT = -1,
	SPR_ERR_CUDA = -2,           /* a CUDA runtime call failed; see sprGetLastError */
	SPR_ERR_NO_DEVICE = -3,      /* no CUDA device / wrong architecture */
	SPR_ERR_OUT_OF_MEMORY = -4,
	SPR_ERR_UNSUPPORTED = -5,
} SprStatus;

/* ------------------------------------------------------------------------ */
/* Math PODs (sf::Vec2/Vec3/Vec4/Quat, src/sf/Vector.h, src/sf/Quaternion.h)  */
/* ------------------------------------------------------------------------ */

typedef struct SprVec2 { float x, y; } SprVec2;
typedef struct SprVec3 { float x, y, z; } SprVec3;
typedef struct SprVec4 { float x, y, z, w; } SprVec4;
typedef struct SprQuat { float x, y, z, w; } SprQuat;

/* sf::Mat34 (src/sf/Matrix.h:86-147): column major, 4 columns of Vec3:
 * v = { m00,m10,m20, m01,m11,m21, m02,m12,m22, m03,m13,m23 }. */
typedef struct SprMat34 { float v[12]; } SprMat34;
/* sf::Mat44: column major 4x4. */
typedef struct SprMat44 { float v[16]; } SprMat44;

/* cl::Transform (src/client/Transform.h:9-17). */
typedef struct SprTransform {
	SprVec3 position;
	SprQuat rotation;
	float scale;
} SprTransform;

/* sf::Bounds3 (src/sf/Geometry.h:38-48). */
typedef struct SprBounds3 { SprVec3 origin, extent; } SprBounds3;

/* sf::Frustum (src/sf/Frustum.h:9-12): 8 planes in SoA form, side[k] / caps[k]
 * hold component k (x,y,z,d) of four planes each. */
typedef struct SprFrustum {
	float side[4][4];
	float caps[4][4];
} SprFrustum;

/* ------------------------------------------------------------------------ */
/* Effect descriptor: flat POD mirror of sv::ParticleSystemComponent          */
/* (src/server/ServerState.h:144-239). Field names and defaults follow the     */
/* reference; sf::Array members become (pointer,count) pairs and the texture   */
/* sf::Symbol becomes an opaque id.                                            */
/* ------------------------------------------------------------------------ */

typedef struct SprBSpline2 {           /* sv::BSpline2, ServerState.h:144-147 */
	const SprVec2 *points;
	uint32_t numPoints;
} SprBSpline2;

typedef struct SprGradientPoint {      /* sv::GradientPoint, :149-153 */
	float t;
	SprVec3 color;
} SprGradientPoint;

typedef struct SprGradient {           /* sv::Gradient, :155-159 */
	SprVec3 defaultColor;              /* default (1,1,1) */
	const SprGradientPoint *points;
	uint32_t numPoints;
} SprGradient;

typedef struct SprRandomSphere {       /* sv::RandomSphere, :161-170 */
	float minTheta, maxTheta;          /* 0, 360 */
	float minPhi, maxPhi;              /* 0, 180 */
	float minRadius, maxRadius;        /* 0, 0 */
	SprVec3 scale;                     /* (1,1,1) */
} SprRandomSphere;

typedef struct SprRandomVec3 {         /* sv::RandomVec3, :172-178 */
	SprVec3 offset;
	SprVec3 boxExtent;
	SprRandomSphere sphere;
	SprVec3 rotation;
} SprRandomVec3;

typedef struct SprGravityPoint {       /* sv::GravityPoint, :180-185 */
	SprVec3 position;
	float radius;                      /* 0.5 */
	float strength;                    /* 1.0 */
} SprGravityPoint;

#define SPR_MAX_GRAVITY_POINTS 16      /* sf::SmallArray<GravityPoint,16>, ParticleSystem.cpp:539 */

typedef struct SprParticleSystemComponent { /* sv::ParticleSystemComponent, :187-239 */
	uint64_t texture;                  /* opaque id replacing sf::Symbol texture */
	int32_t frameCount[2];             /* (1,1) */

	float timeStep;                    /* 0.1 */
	float prewarmTime;                 /* 0 */

	float updateRadius;                /* 10 */
	float cullPadding;                 /* 1.5 */
	uint8_t updateOutOfCamera;         /* false */
	double renderOrder;                /* 0 */

	float spawnTime;                   /* 0.2 */
	float spawnTimeVariance;           /* 0 */
	uint32_t burstAmount;              /* 0 */
	uint32_t burstAmountVariance;      /* 0 (unused by the reference hot path) */
	float emitterOnTime;               /* -1 */
	uint8_t localSpace;                /* false */
	uint8_t instantDelete;             /* false */

	SprRandomVec3 emitPosition;
	SprRandomVec3 emitVelocity;

	SprVec3 emitVelocityAttractorOffset;
	float emitVelocityAttractorStrength; /* NO default in the reference (ServerState.h:212); sprComponentDefaults sets 0 */

	const SprGravityPoint *gravityPoints;
	uint32_t numGravityPoints;         /* <= SPR_MAX_GRAVITY_POINTS */

	float drag;                        /* 0 */
	SprVec3 gravity;                   /* (0,0,0) */

	float size;                        /* 0.5 */
	float sizeVariance;                /* 0 */

	float lifeTime;                    /* 1 */
	float lifeTimeVariance;            /* 0 */

	SprBSpline2 scaleSpline;
	SprBSpline2 alphaSpline;
	SprBSpline2 additiveSpline;
	SprBSpline2 erosionSpline;
	SprGradient gradient;

	float frameRate;                   /* 0 */
	uint8_t relativeFrameRate;         /* false */
	uint8_t randomStartFrame;          /* false */

	float rotation;                    /* 0 */
	float rotationVariance;            /* 0 */
	float spin;                        /* 0 */
	float spinVariance;                /* 0 */
} SprParticleSystemComponent;

/* Fill *c with the reference's default member initialisers (ServerState.h:187-239). */
SPR_API void sprComponentDefaults(SprParticleSystemComponent *c);

/* ------------------------------------------------------------------------ */
/* GPU-side structures the reference renderer consumes                        */
/* ------------------------------------------------------------------------ */

/* ParticleSystemImp::GpuParticle (ParticleSystem.cpp:178-184): one vertex.
 * The vertex stream holds FOUR identical copies per live particle (:607-665),
 * attributes a_PositionLife @0 / a_VelocitySeed @16, stride 32 (:280-281). */
typedef struct SprGpuParticle {
	SprVec3 position;
	float life;
	SprVec3 velocity;
	float seed;
} SprGpuParticle;

/* Particle_VertexType_t (src/game/shader/Particle.h:112-125), 96 bytes. */
typedef struct SprVertexTypeUbo {
	SprVec2 u_FrameCount;
	float u_FrameRate;
	uint8_t _pad_12[4];
	SprVec4 u_RotationControl;
	float u_Additive;
	float u_StartFrame;
	uint8_t _pad_40[8];
	SprVec4 u_SplineMad;
	SprVec2 u_ScaleBaseVariance;
	SprVec2 u_LifeTimeBaseVariance;
	float u_RelativeFrameRate;
	uint8_t _pad_84[12];
} SprVertexTypeUbo;

/* Particle_VertexInstance_t (src/game/shader/Particle.h:100-108), 160 bytes. */
typedef struct SprVertexInstanceUbo {
	float u_ParticleToWorld[16];
	float u_WorldToClip[16];
	SprVec4 u_VisFogMad;
	float u_InvDelta;
	float u_Aspect;
	uint8_t _pad_152[8];
} SprVertexInstanceUbo;

#define SPR_SPLINE_SAMPLE_RATE 64      /* ParticleSystem.cpp:253 */
#define SPR_SPLINE_LUT_BYTES (SPR_SPLINE_SAMPLE_RATE * 4 * 2) /* :326, two RGBA8 rows */

/* ------------------------------------------------------------------------ */
/* Boundary argument PODs                                                     */
/* ------------------------------------------------------------------------ */

/* cl::SystemsDesc (src/client/System.h:28-32); only seed[1] is read
 * (ParticleSystem.cpp:267). */
typedef struct SprSystemsDesc {
	float persistCamera[2];
	float persistZoom;
	uint64_t seed[4];
} SprSystemsDesc;

/* cl::TransformUpdate (src/client/System.h:46-51). */
typedef struct SprTransformUpdate {
	SprTransform previousTransform;
	SprTransform transform;
	SprMat34 entityToWorld;
} SprTransformUpdate;

/* cl::RenderArgs (src/client/System.h:246-256), fields the path reads
 * (ParticleSystem.cpp:840-876). */
typedef struct SprRenderArgs {
	SprVec3 cameraPosition;
	SprMat44 viewToClip;               /* only m00 (v[0]) and m11 (v[5]) are read, :874 */
	SprMat44 worldToClip;
	SprFrustum frustum;
	SprVec4 visFogWorldMad;            /* VisFogImage::worldMad, VisFogSystem.h:11-15, ParticleSystem.cpp:829,:876 */
} SprRenderArgs;

/* cl::FrameArgs (src/client/System.h:258-270), fields the path reads
 * (ParticleSystem.cpp:793, :808-822) plus mainRenderArgs.cameraPosition, which
 * the optional per-particle depth sort uses as its key origin. */
typedef struct SprFrameArgs {
	uint64_t frameIndex;
	double gameTime;
	float dt;
	SprVec3 cameraPosition;            /* FrameArgs::mainRenderArgs.cameraPosition */
} SprFrameArgs;

/* One draw the reference would submit with sg_draw (ParticleSystem.cpp:851-901). */
typedef struct SprDrawRecord {
	uint32_t effectId;
	uint32_t typeId;
	uint32_t numParticles;             /* sg_draw(0, numParticles*6, 1), :901 */
	uint32_t typeUboChanged;           /* 1 when the reference re-applies the type UBO, :880-883 */
	uint64_t vertexByteOffset;         /* offset of the effect's run inside the device vertex stream */
	uint32_t uploadedThisFrame;        /* 1 when the reference would sg_append_buffer the effect's stream in this renderMain
	                                      (its cached copy is older than maxParticleCacheFrames, :861-867) */
	uint32_t uploadRingSlot;           /* uploadFrameIndex % maxParticleCacheFrames: which of the reference's vertex
	                                      buffers holds the effect's newest upload (:869) */
	SprVertexInstanceUbo instanceUbo;  /* :871-876 */
} SprDrawRecord;

typedef struct SprRenderOutput {
	const SprGpuParticle *deviceVertices; /* DEVICE pointer: base of the vertex stream */
	const SprDrawRecord *records;      /* HOST pointer, valid until the next call on the system */
	uint32_t numRecords;
	uint32_t numSortedEffects;         /* size of sortedEffects before the budget cut-off (:848) */
} SprRenderOutput;

/* Parity read-back of one effect (ParticleSystemImp::Effect, ParticleSystem.cpp:186-218). */
typedef struct SprEffectInfo {
	uint32_t typeId;
	uint32_t slotCount;                /* particles.size * 4 */
	uint32_t aliveCount;               /* gpuParticles.size / 4 */
	uint32_t freeCount;                /* freeIndices.size */
	float timeStep;
	float timeDelta;
	float spawnTimer;
	float emitOnTimer;
	uint8_t stopEmit;
	uint8_t firstEmit;
	uint8_t instantDelete;
	uint8_t _pad;
	SprBounds3 gpuBounds;
	double lastUpdateTime;
	uint64_t simSteps;                 /* number of simulateParticlesImp calls so far */
} SprEffectInfo;

/* ------------------------------------------------------------------------ */
/* Context (one per GPU / process)                                            */
/* ------------------------------------------------------------------------ */

typedef struct SprContext SprContext;
typedef struct SprSystem SprSystem;

typedef enum SprContextFlags {
	/* Extension: per-effect back-to-front per-particle
	 * depth sort, key = |cameraPosition - pos|^2 evaluated dx*dx+dy*dy+dz*dz
	 * (src/sf/Vector.h:99), descending, ties by ascending slot, the ordering
	 * idiom of BillboardOrder (src/client/BillboardSystem.cpp:41-51,:119).
	 * Off = strict drop-in: stream in ascending slot order (ParticleSystem.cpp:611-665). */
	SPR_CTX_SORT_PARTICLES = 1u << 0,
} SprContextFlags;

typedef struct SprContextDesc {
	int32_t device;                    /* CUDA device ordinal */
	uint32_t flags;                    /* SprContextFlags */
	uint32_t maxParticlesPerFrame;     /* 0 -> 4096, ParticleSystem.cpp:24 */
	uint32_t maxParticleCacheFrames;   /* 0 -> 16,   ParticleSystem.cpp:25 */
	uint64_t initialSlotCapacity;      /* particle slots reserved up front (0 -> 1 Mi) */
	void *stream;                      /* cudaStream_t to run on; NULL -> context-owned stream */
} SprContextDesc;

SPR_API int sprContextCreate(const SprContextDesc *desc, SprContext **out);
SPR_API int sprContextDestroy(SprContext *ctx);
SPR_API int sprContextSynchronize(SprContext *ctx);
SPR_API void *sprContextStream(SprContext *ctx); /* cudaStream_t all work is enqueued on */
SPR_API const char *sprGetLastError(void);
/* Number of kernel launches this context has enqueued so far (bench `gpu_launches`). */
SPR_API uint64_t sprContextLaunchCount(SprContext *ctx);

/* Totals over every live effect of the context: simulateParticlesImp calls issued so far and
 * the current number of live particles (one device read-back). Either pointer may be NULL. */
SPR_API int sprContextGetTotals(SprContext *ctx, uint64_t *outSimSteps, uint64_t *outAliveParticles);

/* Measurement hooks (no reference counterpart): CUDA events around every kernel launch
 * of the context, accumulated per kernel name. Used by bench.py for the roofline. */
SPR_API int sprContextEnableKernelTiming(SprContext *ctx, int enable);
SPR_API int sprContextGetKernelTimes(SprContext *ctx, int reset, const char **names, double *ms, uint64_t *launches,
	uint32_t capacity, uint32_t *outCount);

/* ------------------------------------------------------------------------ */
/* System API == cl::ParticleSystem (src/client/ParticleSystem.h:9-22)         */
/* ------------------------------------------------------------------------ */

/* ParticleSystem::create(const SystemsDesc&), ParticleSystem.h:11, ParticleSystem.cpp:265-267,:907 */
SPR_API int sprSystemCreate(SprContext *ctx, const SprSystemsDesc *desc, SprSystem **out);
SPR_API int sprSystemDestroy(SprSystem *sys);

/* reserveEffectType, ParticleSystem.h:13, ParticleSystem.cpp:696-720 */
SPR_API int sprReserveEffectType(SprSystem *sys, const SprParticleSystemComponent *c, uint32_t *outTypeId);
/* releaseEffectType, ParticleSystem.h:14, ParticleSystem.cpp:722-728 */
SPR_API int sprReleaseEffectType(SprSystem *sys, const SprParticleSystemComponent *c);

/* addEffect, ParticleSystem.h:16, ParticleSystem.cpp:730-766. The effect id the
 * reference hands to Entities::addComponent (:765) is returned in *outEffectId. */
SPR_API int sprAddEffect(SprSystem *sys, uint32_t entityId, uint8_t componentIndex,
	const SprParticleSystemComponent *c, const SprTransform *transform, uint32_t *outEffectId);

/* EntitySystem::updateTransform, System.h:60, ParticleSystem.cpp:768-784 */
SPR_API int sprUpdateTransform(SprSystem *sys, uint32_t effectId, const SprTransformUpdate *update);
/* EntitySystem::prepareForRemove, System.h:61, ParticleSystem.cpp:786-794 */
SPR_API int sprPrepareForRemove(SprSystem *sys, uint32_t effectId, const SprFrameArgs *args, int *outCanRemove);
/* EntitySystem::remove, System.h:62, ParticleSystem.cpp:796-806 */
SPR_API int sprRemoveEffect(SprSystem *sys, uint32_t effectId);

/* updateParticles(const VisibleAreas &active, const FrameArgs&), ParticleSystem.h:18,
 * ParticleSystem.cpp:808-822. activeIds = active.get(AreaGroup::ParticleEffect).
 * Every id must name a live effect (SPR_ERR_INVALID_ARGUMENT otherwise). An id listed twice in one list is refused with
 * SPR_ERR_UNSUPPORTED before anything is changed: the reference would run that effect's sub-step loop once per entry
 * (:812-821), the device plan is per effect. */
SPR_API int sprUpdateParticles(SprSystem *sys, const uint32_t *activeIds, uint32_t numActive,
	const SprFrameArgs *args);

/* Same call for many independent systems of one context in ONE set of kernel
 * launches (each system keeps its own RNG and id spaces, ParticleSystem.cpp:241-251).
 * Equivalent to calling sprUpdateParticles(systems[i], activeIds[i], numActive[i], &args[i])
 * for i = 0..numSystems-1. If args is shared, pass argsStride = 0.
 * A system listed twice is refused (SPR_ERR_UNSUPPORTED: its effects share one RNG stream, merge the lists); ids are
 * checked as in sprUpdateParticles. dt, gameTime and -- with SPR_CTX_SORT_PARTICLES -- the camera of the depth keys are
 * per system: every system is sorted towards its own args[i].cameraPosition. */
SPR_API int sprUpdateParticlesBatch(SprContext *ctx, SprSystem *const *systems, uint32_t numSystems,
	const uint32_t *const *activeIds, const uint32_t *numActive,
	const SprFrameArgs *args, size_t argsStride);

/* renderMain(const VisFogSystem*, const VisibleAreas &visible, const RenderArgs&),
 * ParticleSystem.h:20, ParticleSystem.cpp:824-903. */
SPR_API int sprRenderMain(SprSystem *sys, const uint32_t *visibleIds, uint32_t numVisible,
	const SprRenderArgs *args, SprRenderOutput *out);

/* renderMain for many systems of one context with a single device read-back. Equivalent to
 * sprRenderMain(systems[i], visibleIds[i], numVisible[i], &args[i], &outs[i]) for every i;
 * pass argsStride = 0 when the RenderArgs are shared. A batch of at least 32768 visible effects in at least 64
 * systems builds its draw records on a few helper threads (joined before the call returns; SPR_RENDER_THREADS=1 in the
 * environment keeps everything on the calling thread, and so does a batch that names a system twice). */
SPR_API int sprRenderMainBatch(SprContext *ctx, SprSystem *const *systems, uint32_t numSystems,
	const uint32_t *const *visibleIds, const uint32_t *numVisible,
	const SprRenderArgs *args, size_t argsStride, SprRenderOutput *outs);

/* ------------------------------------------------------------------------ */
/* Parity read-backs (host copies; each synchronises the context stream)      */
/* ------------------------------------------------------------------------ */

SPR_API int sprGetEffectInfo(SprSystem *sys, uint32_t effectId, SprEffectInfo *out);
/* freeIndices in stack order (ParticleSystem.cpp:217). out has room for `capacity` entries. */
SPR_API int sprReadFreeIndices(SprSystem *sys, uint32_t effectId, uint32_t *out, uint32_t capacity, uint32_t *outCount);
/* The effect's vertex stream (4 copies per particle) exactly as gpuParticles (:216). */
SPR_API int sprReadVertexStream(SprSystem *sys, uint32_t effectId, SprGpuParticle *out, uint32_t capacityVertices, uint32_t *outCount);
/* Slot state {pos,life,vel,seed} for slots [0,slotCount) (Particle4 lanes, :170-176). */
SPR_API int sprReadSlots(SprSystem *sys, uint32_t effectId, SprGpuParticle *out, uint32_t capacitySlots, uint32_t *outCount);
/* Slot index of the k-th particle of the current stream (compaction / sort order). */
SPR_API int sprReadStreamSlots(SprSystem *sys, uint32_t effectId, uint32_t *out, uint32_t capacity, uint32_t *outCount);
/* updateCtx.rng (state, inc), ParticleSystem.cpp:251. */
SPR_API int sprGetRngState(SprSystem *sys, uint64_t outStateInc[2]);
/* Per-type LUT (512 B, :326-400) and type UBO (:312-321,:410-419). Byte 3 of each
 * gradient texel is never written by the reference (:398-399); this library writes 0. */
SPR_API int sprGetEffectTypeLut(SprSystem *sys, uint32_t typeId, uint8_t out[SPR_SPLINE_LUT_BYTES]);
SPR_API int sprGetEffectTypeUbo(SprSystem *sys, uint32_t typeId, SprVertexTypeUbo *out);

/* Host-only type bake (initEffectType, ParticleSystem.cpp:298-450) without a context:
 * the 512-byte spline/gradient LUT and the 96-byte type UBO for `typeId`. Needs no GPU. */
SPR_API int sprBakeEffectType(const SprParticleSystemComponent *c, uint32_t typeId,
	uint8_t outLut[SPR_SPLINE_LUT_BYTES], SprVertexTypeUbo *outUbo);

/* ------------------------------------------------------------------------ */
/* Area query                                         */
/* ------------------------------------------------------------------------ */

/* The effects of `sys` whose sphere area intersects `frustum`: what cl::AreaSystem::queryFrustum
 * (src/client/AreaSystem.cpp:139-170, sf::Frustum::intersects(Sphere), src/sf/Frustum.h:42-58) returns for the
 * AreaGroup::ParticleEffect spheres the particle system registers -- centre = the effect's transform position at
 * addEffect / the last updateTransform, radius = the descriptor's updateRadius (ParticleSystem.cpp:762-763,
 * :782-783, removed at :801). The ids come back in ascending order (the reference's order is that of its spatial
 * tree); these are the lists sprUpdateParticles / sprRenderMain take. outCount is the full count even when it
 * exceeds `capacity`. Synchronises the context stream. */
SPR_API int sprQueryEffects(SprSystem *sys, const SprFrustum *frustum, uint32_t *outIds, uint32_t capacity, uint32_t *outCount);

/* ------------------------------------------------------------------------ */
/* Descriptors from the reference's on-disk format    */
/* ------------------------------------------------------------------------ */

typedef struct SprDescriptorSet SprDescriptorSet;

/* Reads effect descriptors from JSON in the format sp::writeJson produces over the reflection table of
 * sv::ParticleSystemComponent (src/server/ServerStateReflection.cpp:321-363) with the lenient dialect of
 * loadConfig (src/server/ServerState.cpp:786-796: bare keys, comments, missing / trailing commas): a prefab
 * ({ components: [ { type: "ParticleSystem", .. }, .. ] }, sv::Prefab, ServerState.h:568-572), one tagged
 * component, or a bare component object. sp::readJson's rules apply (src/sp/Json.cpp:116-270): absent fields keep
 * their defaults, unknown keys are ignored, a value of the wrong type or an unknown component type fails the load
 * (SPR_ERR_INVALID_ARGUMENT, text in sprGetLastError). Host only, no GPU needed. The set owns the descriptors and
 * their arrays; their addresses are stable, so they can be passed to sprReserveEffectType / sprAddEffect. */
SPR_API int sprDescriptorsFromJson(const char *json, size_t length, SprDescriptorSet **out);
SPR_API uint32_t sprDescriptorCount(const SprDescriptorSet *set);
SPR_API const SprParticleSystemComponent *sprDescriptorGet(const SprDescriptorSet *set, uint32_t index);
/* sf::Symbol texture of descriptor `index` ("" when absent); SprParticleSystemComponent::texture is its FNV-1a 64 hash. */
SPR_API const char *sprDescriptorTextureName(const SprDescriptorSet *set, uint32_t index);
SPR_API void sprDescriptorsFree(SprDescriptorSet *set);

/* ------------------------------------------------------------------------ */
/* Vertex stage                                       */
/* ------------------------------------------------------------------------ */

/* What the reference's vertex shader writes for one vertex
 * (src/game/shader/Particle.glsl:73-140: gl_Position, v_Uv0, v_Uv1, flat v_Color, flat v_Params). */
typedef struct SprShadedVertex {
	float position[4];
	float uv0[2];
	float uv1[2];
	float color[4];
	float params[4];
} SprShadedVertex;

/* u_VisFogTexture (Particle.glsl:13, :43-50): RG8 unorm texels, row major, bilinear,
 * clamp-to-edge. A DEVICE pointer; the image belongs to the caller (cl::VisFogSystem is outside the path). */
typedef struct SprFogImage {
	const void *deviceTexels;
	uint32_t width, height;
} SprFogImage;

/* Runs Particle.glsl `vs` (src/game/shader/Particle.glsl:73-140) over the draw records sprRenderMain
 * returned (ParticleSystem.cpp:851-901): record r contributes 4 * numParticles vertices, in record order,
 * particle k of it at vertices 4k..4k+3 (corner = vertex & 3, Particle.glsl:91), written to the DEVICE
 * buffer `deviceOut` (capacityVertices entries; records that do not fit are cut off at a quad boundary).
 * `fog` may be NULL: every particle fully visible. Stream-ordered on the context's stream, not
 * synchronised. Texture filtering is done in float (hardware quantises the weights to 8 bits). */
SPR_API int sprShadeVertices(SprSystem *sys, const SprDrawRecord *records, uint32_t numRecords, const SprFogImage *fog,
	SprShadedVertex *deviceOut, uint64_t capacityVertices, uint64_t *outNumVertices);

/* ------------------------------------------------------------------------ */
/* Billboards: cl::BillboardSystem                    */
/* ------------------------------------------------------------------------ */

typedef struct SprBillboardSystem SprBillboardSystem;

/* The fields of a loaded sp::Sprite that BillboardSystem::renderMain reads (src/sp/Sprite.h:44-57,
 * src/client/BillboardSystem.cpp:126-143). atlas is an opaque id (the sp::Atlas pointer); 0 = sprite
 * missing or not loaded: addBillboard drops the billboard (:82, :100). */
typedef struct SprSprite {
	uint64_t atlas;
	SprVec2 minVert, maxVert;
	SprVec2 vertUvScale, vertUvBias;
} SprSprite;

/* cl::Billboard (src/client/BillboardSystem.h:9-18). transform is sf::Mat34: column major,
 * m00 m10 m20 | m01 m11 m21 | m02 m12 m22 | m03 m13 m23 (src/sf/Matrix.h:86-97). */
typedef struct SprBillboard {
	SprSprite sprite;
	float transform[12];
	SprVec4 color;
	float depth;
	SprVec2 anchor, cropMin, cropMax;
} SprBillboard;

/* BillboardVertex (src/client/BillboardSystem.cpp:53-58), 24 bytes. */
typedef struct SprBillboardVertex {
	float position[3];
	float texCoord[2];
	uint32_t color;                    /* packColor(), :12-28 */
} SprBillboardVertex;

/* One sg_draw of renderMain (:198-203, :213-225): `count` quads of one atlas from quad `firstQuad` on. */
typedef struct SprBillboardDraw {
	uint64_t atlas;
	uint32_t count;
	uint32_t firstQuad;
} SprBillboardDraw;

typedef struct SprBillboardOutput {
	const SprBillboardVertex *deviceVertices; /* DEVICE pointer: 4 vertices per quad, sorted (:163-196) */
	uint32_t numQuads;
	uint32_t numDraws;
	const SprBillboardDraw *draws;     /* HOST pointer, valid until the next call on the system */
	float worldToClip[16];             /* Billboard_Vertex_t (:208-210) */
} SprBillboardOutput;

/* BillboardSystem::create (src/client/BillboardSystem.h:22). maxBillboardsPerFrame 0 = the reference's 4096 (:10). */
SPR_API int sprBillboardSystemCreate(SprContext *ctx, uint32_t maxBillboardsPerFrame, SprBillboardSystem **out);
SPR_API int sprBillboardSystemDestroy(SprBillboardSystem *sys);
/* addBillboard (:80-117), `count` times in array order; the 4-argument overload of the reference is the
 * same call with anchor 0.5, crop [0, 1]. Host only: the frame's billboards go to the device in renderMain. */
SPR_API int sprBillboardAdd(SprBillboardSystem *sys, const SprBillboard *billboards, uint32_t count);
/* The same billboard in the form the device reads (120 bytes): what sprBillboardAdd makes of an SprBillboard with a
 * loaded sprite -- the sprite's fields, the transform, the colour packed by packColor() (:12-28), anchor, crop, depth. */
typedef struct SprBillboardPacked {
	uint64_t atlas;                    /* sprite->atlas, not 0 */
	float minVert[2], maxVert[2], vertUvScale[2], vertUvBias[2];
	float transform[12];               /* sf::Mat34, column major */
	float anchor[2];
	uint32_t color;
	float cropMin[2], cropMax[2];
	float depth;
} SprBillboardPacked;
/* Host-side conversion (no system, no device): packs `count` billboards in order, skipping those without a loaded
 * sprite like addBillboard does (:82, :100); *outCount = entries written to `out` (room for `count`). */
SPR_API int sprBillboardPack(const SprBillboard *billboards, uint32_t count, SprBillboardPacked *out, uint32_t *outCount);
#define SPR_MEMORY_HOST 0
#define SPR_MEMORY_DEVICE 1
/* addBillboard for billboards that are already packed, `count` times in array order, in host memory (pinned memory
 * copies at the rate of the bus; no per-billboard host work) or in device memory of the context's GPU (a producer on the
 * GPU, or billboards that do not change between frames: nothing crosses the bus). May be mixed with sprBillboardAdd:
 * the frame's billboards are those of all calls in call order. The array is read by sprBillboardRenderMain: it must
 * stay valid and unchanged until that call returns. */
SPR_API int sprBillboardAddPacked(SprBillboardSystem *sys, const SprBillboardPacked *packed, uint32_t count, int memory);
/* renderMain (:119-228): sort by (depth, insertion index), cut to the frame budget, crop, expand;
 * clears the frame's billboards. Synchronises the context stream (the draw list is host data). */
SPR_API int sprBillboardRenderMain(SprBillboardSystem *sys, const SprRenderArgs *args, SprBillboardOutput *out);
/* Parity read-back: host copy of the vertex buffer of the last renderMain (billboardVertices, :65). */
SPR_API int sprBillboardReadVertices(SprBillboardSystem *sys, SprBillboardVertex *out, uint32_t capacityQuads, uint32_t *outQuads);

/* Device pointers for zero-copy consumers (benchmarks, the multi-GPU gather).
 * LIFETIME of every device pointer and offset this API hands out (SprRenderOutput::deviceVertices,
 * SprDrawRecord::vertexByteOffset, the fields below): valid until the next sprAddEffect or sprUpdateParticles[Batch] on
 * the context. Those calls may grow the slot pool (all pool arrays are re-allocated) or move an effect that outgrew its
 * slot range (its slotBase and with it its vertex offset change). A context created with initialSlotCapacity large
 * enough for everything it will hold never re-allocates the pool; an effect whose particle count is bounded by its
 * descriptor (burst + 20 spawns per step over its lifetime) never moves. */
typedef struct SprEffectDeviceView {
	const SprGpuParticle *vertices;    /* DEVICE: effect's run, 4 vertices per particle */
	const SprGpuParticle *slots;       /* DEVICE: slot state */
	const uint32_t *aliveCount;        /* DEVICE: live particle count of the last step */
	uint64_t vertexByteOffset;
	uint32_t slotCapacity;
} SprEffectDeviceView;
SPR_API int sprGetEffectDeviceView(SprSystem *sys, uint32_t effectId, SprEffectDeviceView *out);

/* ------------------------------------------------------------------------ */
/* Multi-GPU draw-stream assembly (BASELINE.json north_star: gather of        */
/* per-GPU sorted runs). Packs the 32-byte record of every live particle of    */
/* the listed effects, in stream order, into one contiguous DEVICE buffer      */
/* (the run this GPU contributes), and expands gathered runs 4x afterwards.    */
/* ------------------------------------------------------------------------ */

SPR_API int sprPackRuns(SprContext *ctx, SprSystem *const *systems, const uint32_t *effectIds, uint32_t numEffects,
	SprGpuParticle *deviceOut, uint64_t capacityRecords, uint32_t *deviceCounts /* numEffects */, uint64_t *outTotalRecords /* host, may be NULL (no sync) */);
SPR_API int sprExpandRuns(SprContext *ctx, const SprGpuParticle *deviceRecords, uint64_t numRecords,
	SprGpuParticle *deviceVerticesOut);

/* Fused gather + expansion over NVLink peer memory (one process per GPU).
 *
 * Each rank packs its sorted 32-byte runs into a run buffer that the other ranks map through
 * CUDA IPC; ONE kernel per rank then reads every rank's run straight out of peer memory
 * (plain loads on the mapped pointers, the transfer overlapping the 4x expansion tile by tile)
 * and writes the whole draw stream locally. No intermediate gathered-records buffer, no NCCL copy.
 * The caller provides the cross-rank ordering: a stream-ordered barrier between sprRunBufferPack on
 * every rank and sprExpandFromPeers (bench.py uses a 4-byte NCCL all_reduce), and double-buffers
 * the run buffers so that the next step's pack never overwrites a buffer a peer may still read. */
typedef struct SprRunBuffer SprRunBuffer;
#define SPR_IPC_HANDLE_BYTES 64
SPR_API int sprRunBufferCreate(SprContext *ctx, uint64_t capacityRecords, SprRunBuffer **out);
SPR_API int sprRunBufferDestroy(SprRunBuffer *buf);
SPR_API int sprRunBufferGetIpcHandle(SprRunBuffer *buf, uint8_t outHandle[SPR_IPC_HANDLE_BYTES]);
/* Map a peer's run buffer (same node). Returns a device pointer valid in this process. */
SPR_API int sprRunBufferOpenPeer(SprContext *ctx, const uint8_t handle[SPR_IPC_HANDLE_BYTES], void **outDevicePtr);
SPR_API int sprRunBufferClosePeer(SprContext *ctx, void *devicePtr);
SPR_API void *sprRunBufferDevicePtr(SprRunBuffer *buf);
/* sprPackRuns into `buf` (record count stored in the buffer header, no host sync). */
SPR_API int sprRunBufferPack(SprContext *ctx, SprSystem *const *systems, const uint32_t *effectIds, uint32_t numEffects, SprRunBuffer *buf);
/* bufferPtrs[r] = run buffer of rank r (own pointer for the own rank, mapped pointers for peers), rank-major
 * output: deviceVerticesOut receives 4 vertices per record of rank 0's run, then rank 1's, ...
 * deviceTotalRecords (may be NULL) receives the total record count (device uint64). myRank only staggers the
 * order in which the peers are read (every GPU starts on a different peer and reads all of them concurrently). */
SPR_API int sprExpandFromPeers(SprContext *ctx, void *const *bufferPtrs, uint32_t numRanks, uint32_t myRank,
	SprGpuParticle *deviceVerticesOut, uint64_t capacityRecords, uint64_t *deviceTotalRecords);
/* How much of the GPU sprExpandFromPeers takes: resident CTAs (256 threads) per SM of its persistent grid, 1..8
 * (0 = default 8, the whole GPU). The assembly is bound by NVLink (all GPUs pulling from all: ~650 GB/s into each GPU
 * whether 2 or 8 CTAs per SM pull, profiles/r02_peer_allgather_microbench_8gpu.txt), so a caller that runs it on a second
 * context / stream under the next frame's simulation gives it a small share and leaves the rest of every SM to the
 * simulation kernels (bench.py: gather.fused_overlapped). */
/* At a share of 1 or 2 the assembly runs as a bulk-asynchronous pipeline (cp.async.bulk loads out of the peers into shared
 * memory, bulk tensor stores of the 4x expansion: no register holds data in flight, a third of the registers of the
 * load/store kernel), which leaves the simulation beside it more of the SM; same stream, bit for bit. */
SPR_API int sprContextSetAssemblyShare(SprContext *ctx, uint32_t ctasPerSm);
/* Which kernel sprExpandFromPeers runs: 0 = chosen by the share (bulk-asynchronous at 1 or 2 CTAs per SM, load/store above),
 * 1 = the load/store kernel, 2 = the bulk-asynchronous pipeline. Same stream either way; which one is faster under a given
 * simulation depends on what bounds the assembly (NVLink ingress at 8 GPUs, HBM stores at 2-4): bench.py tries both. */
SPR_API int sprContextSetAssemblyKernel(SprContext *ctx, uint32_t kernel);
/* Deferred expansion (SPR_CTX_SORT_PARTICLES contexts only, else SPR_ERR_UNSUPPORTED), for pipelines whose consumer is the
 * ASSEMBLED draw stream (sprPackRuns / sprRunBufferPack -> sprExpandRuns / sprExpandFromPeers), which contains this
 * GPU's shard too: while on, sprUpdateParticles[Batch] emits, integrates, compacts and sorts as always but does not write
 * the local 4x vertex stream, and the pack calls gather the sorted 32-byte records from the slot state (same bytes in
 * the run). The local stream -- SprDrawRecord::vertexByteOffset, sprReadVertexStream, sprShadeVertices -- is stale for
 * effects updated while it is on (effects that fit one 128-slot granule still write theirs); everything else (counts,
 * bounds, free lists, sprReadStreamSlots) is as always. Saves the 128 bytes per particle the local expansion writes
 * and the strided read of the pack: bench.py gather.fused_deferred. */
SPR_API int sprContextSetDeferredExpansion(SprContext *ctx, int on);

#ifdef __cplusplus
}
#endif

#endif /* SPEAR_PARTICLES_H */
