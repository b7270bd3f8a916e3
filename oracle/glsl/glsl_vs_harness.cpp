// glsl_vs_harness.cpp -- TEST INFRASTRUCTURE ONLY: runs the reference's shipped particle vertex shader
// (Particle_vs_source_glsl330, /root/reference/src/game/shader/Particle.h:239, generated from Particle.glsl:73-140)
// over particle records on the CPU. particle_vs.gen.inc is NOT in this repository: oracle/glsl/build_glsl_vs.py
// decodes the shader text out of the reference tree into a temporary directory at build time, rewrites only its
// interface declarations, and this file compiles the body unchanged against oracle/glsl/glsl_shim.h.
// Output: oracle/_ref/libspear_ref_glslvs.so (git-ignored; travels to the GPU box like the other reference builds).
#include <stdint.h>
#include <string.h>

#include "glsl_shim.h"

namespace glsl {
static int gl_VertexID;
static vec4 gl_Position;
#include "particle_vs.gen.inc" // -> VertexInstance[10], VertexType[6], samplers, attributes, varyings, glsl_main()
} // namespace glsl

extern "C" {

// the decoded GLSL 3.30 text this library was built from (provenance, shown by the tests)
__attribute__((visibility("default"))) const char *glslvs_source(void) { return glsl::kShaderSource; }

// records: numParticles x {pos.xyz, life, vel.xyz, seed}; instanceUbo: 40 floats (Particle_VertexInstance_t);
// typeUbo: 24 floats (Particle_VertexType_t); splineAtlas: RGBA8 atlasW x atlasH, GL_REPEAT (sokol's default wrap,
// ParticleSystem.cpp:285-294 sets none); fog: RG8 fogW x fogH, CLAMP_TO_EDGE (VisFogSystem.cpp:70-77) or NULL = a
// fully visible 1 x 1 image (factor (1, 1): the colour is multiplied by exactly 1).
// out: 4 * numParticles x {gl_Position, v_Uv0, v_Uv1, v_Color, v_Params} = 16 floats (SprShadedVertex layout);
// vertex 4 * p + c runs with gl_VertexID = 4 * p + c (ParticleSystem.cpp:892-899: 4 vertices per particle).
__attribute__((visibility("default"))) int glslvs_shade(const float *records, uint32_t numParticles, const float *instanceUbo, const float *typeUbo,
	const uint8_t *splineAtlas, uint32_t atlasW, uint32_t atlasH, const uint8_t *fog, uint32_t fogW, uint32_t fogH, float *out)
{
	using namespace glsl;
	static const uint8_t clearFog[2] = { 255, 255 };
	if (!records || !instanceUbo || !typeUbo || !splineAtlas || !out) return -1;
	memcpy(VertexInstance, instanceUbo, sizeof(float) * 40);
	memcpy(VertexType, typeUbo, sizeof(float) * 24);
	u_SplineTexture.texels = splineAtlas; u_SplineTexture.width = (int)atlasW; u_SplineTexture.height = (int)atlasH;
	u_SplineTexture.channels = 4; u_SplineTexture.repeat = true;
	u_VisFogTexture.texels = fog ? fog : clearFog; u_VisFogTexture.width = fog ? (int)fogW : 1; u_VisFogTexture.height = fog ? (int)fogH : 1;
	u_VisFogTexture.channels = 2; u_VisFogTexture.repeat = false;
	for (uint32_t p = 0; p < numParticles; p++) {
		const float *r = records + (size_t)p * 8;
		for (uint32_t c = 0; c < 4; c++) {
			a_PositionLife = vec4(r[0], r[1], r[2], r[3]);
			a_VelocitySeed = vec4(r[4], r[5], r[6], r[7]);
			gl_VertexID = (int)(4 * p + c);
			glsl_main();
			float *o = out + ((size_t)p * 4 + c) * 16;
			o[0] = gl_Position.x; o[1] = gl_Position.y; o[2] = gl_Position.z; o[3] = gl_Position.w;
			o[4] = v_Uv0.x; o[5] = v_Uv0.y; o[6] = v_Uv1.x; o[7] = v_Uv1.y;
			o[8] = v_Color.x; o[9] = v_Color.y; o[10] = v_Color.z; o[11] = v_Color.w;
			o[12] = v_Params.x; o[13] = v_Params.y; o[14] = v_Params.z; o[15] = v_Params.w;
		}
	}
	return 0;
}

} // extern "C"
