/*
 * oracle_api.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Uniform C API implemented twice:
 *   oracle/_ref/libspear_ref.so   kind "reference": the reference's own
 *       src/client/ParticleSystem.cpp compiled headless (oracle/ref_harness.cpp,
 *       recipe oracle/build_ref.py);
 *   oracle/libspear_oracle.so     kind "port": a plain-C restatement of the same
 *       algorithm (oracle/spear_oracle.c).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load these libraries, and only as the checker / the CPU
 * baseline. The product (spear_b200/, include/spear_particles.h) never links,
 * loads or calls anything under oracle/.
 *
 * The POD argument types are shared with the product ABI on purpose, so a
 * parity test feeds byte-identical descriptors to both sides.
 */
#ifndef SPEAR_ORACLE_API_H
#define SPEAR_ORACLE_API_H

#include "../include/spear_particles.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API const char *orc_kind(void);                       /* "reference" or "port" */
ORC_API void *orc_create(const SprSystemsDesc *desc);     /* ParticleSystem::create */
ORC_API void orc_destroy(void *sys);
ORC_API uint32_t orc_reserve_type(void *sys, const SprParticleSystemComponent *c);
ORC_API void orc_release_type(void *sys, const SprParticleSystemComponent *c);
ORC_API uint32_t orc_add_effect(void *sys, uint32_t entityId, uint8_t componentIndex,
	const SprParticleSystemComponent *c, const SprTransform *transform);
ORC_API void orc_update_transform(void *sys, uint32_t effectId, const SprTransformUpdate *update);
ORC_API int orc_prepare_for_remove(void *sys, uint32_t effectId, const SprFrameArgs *args);
ORC_API void orc_remove(void *sys, uint32_t effectId);
ORC_API void orc_update(void *sys, const uint32_t *activeIds, uint32_t numActive, const SprFrameArgs *args);
/* Returns number of draw records written (<= capacity); *numSorted = sortedEffects.size. */
ORC_API uint32_t orc_render(void *sys, const uint32_t *visibleIds, uint32_t numVisible,
	const SprRenderArgs *args, SprDrawRecord *out, uint32_t capacity, uint32_t *numSorted);

/* What the last orc_render handed to sokol (reference harness and adapter harness only; the plain-C port has no
 * renderer side): one entry per sg_draw with the bindings applied for it, and one per sg_append_buffer call. */
typedef struct OrcDrawCapture { uint32_t vertexBufferRing, vertexOffset, indexBuffer, vsSplineImage, vsVisFogImage, fsTextureImage, numElements, _pad; } OrcDrawCapture;
typedef struct OrcAppendCapture { uint32_t bufferRing, offset, numBytes, _pad; uint64_t fnv1a64; } OrcAppendCapture;
ORC_API uint32_t orc_render_capture(void *sys, OrcDrawCapture *draws, uint32_t capDraws, OrcAppendCapture *appends, uint32_t capAppends, uint32_t *numAppends);
/* The 512 x 512 RGBA8 spline atlas as the sg_make_image / sg_bqq_copy_subimages calls left it; returns 0 if untouched. */
ORC_API int orc_read_atlas(void *sys, uint8_t *out);
/* Sphere areas registered with the harness's area system: 5 floats each {effect id, x, y, z, radius}; returns the count. */
ORC_API uint32_t orc_area_spheres(void *sys, float *out, uint32_t capacity);

ORC_API void orc_effect_info(void *sys, uint32_t effectId, SprEffectInfo *out);
ORC_API uint32_t orc_read_free(void *sys, uint32_t effectId, uint32_t *out, uint32_t capacity);
ORC_API uint32_t orc_read_stream(void *sys, uint32_t effectId, SprGpuParticle *out, uint32_t capacityVertices);
ORC_API uint32_t orc_read_slots(void *sys, uint32_t effectId, SprGpuParticle *out, uint32_t capacitySlots);
ORC_API void orc_rng_state(void *sys, uint64_t outStateInc[2]);
ORC_API void orc_type_lut(void *sys, uint32_t typeId, uint8_t out[SPR_SPLINE_LUT_BYTES]);
ORC_API void orc_type_ubo(void *sys, uint32_t typeId, SprVertexTypeUbo *out);

/* Build an sf::Frustum the way the reference does (src/sf/Frustum.h:16-40). */
ORC_API void orc_make_frustum(const SprMat44 *worldToClip, float clipNearW, SprFrustum *out);

/* Timing helper for the CPU baseline: run `steps` updates of all listed effects
 * with dt = args->dt and return elapsed seconds (steady clock); *particleSteps
 * accumulates sum over steps of alive counts. */
ORC_API double orc_bench_steps(void *sys, const uint32_t *activeIds, uint32_t numActive,
	const SprFrameArgs *args, uint32_t steps, uint64_t *particleSteps);

#ifdef __cplusplus
}
#endif
#endif
