// vertex_stage.cu -- SURVEY.md 8(f) row N2: the reference's particle vertex shader
// (src/game/shader/Particle.glsl:73-140, `vs main`) as a CUDA kernel over the device vertex stream.
//
// The hot path writes the stream the reference renderer consumes: 4 identical 32-byte records per
// particle (ParticleSystem.cpp:607-665); the corner expansion, curve / gradient look-up, rotation,
// flip-book frame and fog tint happen in the vertex shader. This kernel produces what that shader
// outputs, per vertex: gl_Position, v_Uv0, v_Uv1 and the flat v_Color / v_Params (64 bytes), for
// consumers that want finished clip-space quads instead of running the shader themselves.
//
//   one thread per particle (the four vertices of a quad share everything but the corner),
//   one CTA per 256 particles of ONE draw record; the 64 x 2 RGBA8 LUT rows of the record's effect
//   type (ParticleSystem.cpp:326-400) are staged in shared memory, the type and instance UBOs
//   (:312-321, :410-419, :871-876) come from the job table.
//
// Texture filtering is restated in float: the spline atlas is sampled with SG_FILTER_LINEAR
// (:287-288) at u = clamp(lifeTime) * 63/512 + (x + 0.5)/512 (:403-413), i.e. at texel coordinate
// x + 63 * clamp(lifeTime) of the type's row: lerp of texels floor and floor + 1. Graphics hardware
// quantises the lerp weight to 8 bits; this kernel (and its oracle, oracle/vertex_stage.py) do not.

#include <cuda_runtime.h>
#include <stdint.h>

#include "device_types.h"
#include "kernels.h"

namespace spr {

__device__ __forceinline__ float fractf(float x) { return x - floorf(x); }                 // GLSL fract
__device__ __forceinline__ float modf_glsl(float x, float y) { return x - y * floorf(x / y); } // GLSL mod
__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

__device__ __forceinline__ float srgb_to_linear(float x) // Particle.glsl:59-66
{
	x = clamp01(x);
	if (x <= 0.04045f) return 0.0773993808f * x;
	return powf(x * 0.947867298578f + 0.0521327014218f, 2.4f);
}

__device__ __forceinline__ void st256(float4 *p, const float4 &a, const float4 &b)
{
	asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
		:: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// lerp of two RGBA8 texels, channel c, as unorm floats
__device__ __forceinline__ float lut_channel(uint32_t t0, uint32_t t1, int c, float f)
{
	float a = (float)((t0 >> (8 * c)) & 255u) * (1.0f / 255.0f);
	float b = (float)((t1 >> (8 * c)) & 255u) * (1.0f / 255.0f);
	return a + (b - a) * f;
}

// bilinear fetch of an RG8 image with clamp-to-edge addressing; returns (r, g) as unorm floats
__device__ inline float2 fog_fetch(const uint8_t *texels, uint32_t w, uint32_t h, float u, float v)
{
	float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
	float fx0 = floorf(x), fy0 = floorf(y);
	float wx = x - fx0, wy = y - fy0;
	int x0 = (int)fx0, y0 = (int)fy0;
	int x1 = x0 + 1, y1 = y0 + 1;
	x0 = min(max(x0, 0), (int)w - 1); x1 = min(max(x1, 0), (int)w - 1);
	y0 = min(max(y0, 0), (int)h - 1); y1 = min(max(y1, 0), (int)h - 1);
	float2 r;
	float c[2];
#pragma unroll
	for (int k = 0; k < 2; k++) {
		float t00 = (float)texels[2 * ((size_t)y0 * w + x0) + k] * (1.0f / 255.0f);
		float t10 = (float)texels[2 * ((size_t)y0 * w + x1) + k] * (1.0f / 255.0f);
		float t01 = (float)texels[2 * ((size_t)y1 * w + x0) + k] * (1.0f / 255.0f);
		float t11 = (float)texels[2 * ((size_t)y1 * w + x1) + k] * (1.0f / 255.0f);
		float top = t00 + (t10 - t00) * wx, bot = t01 + (t11 - t01) * wx;
		c[k] = top + (bot - top) * wy;
	}
	r.x = c[0]; r.y = c[1];
	return r;
}

__global__ void __launch_bounds__(256) k_shade_vertices(const float4 *vertices, const ShadeJob *jobs, uint32_t numJobs,
	const uint32_t *luts, const uint8_t *fogTexels, uint32_t fogW, uint32_t fogH, float4 *out, uint64_t capacityVertices)
{
	__shared__ uint32_t sLut[128];
	__shared__ ShadeJob sJob;
	__shared__ __align__(16) float sStage[8][32 * 36]; // per warp: 32 lanes x 128 bytes at a 144-byte pitch
	// job of this CTA: jobs are ordered by firstBlock
	uint32_t lo = 0, hi = numJobs;
	while (hi - lo > 1) {
		uint32_t mid = (lo + hi) >> 1;
		if (jobs[mid].firstBlock <= blockIdx.x) lo = mid; else hi = mid;
	}
	{
		const uint32_t *src = (const uint32_t*)&jobs[lo];
		uint32_t *dst = (uint32_t*)&sJob;
		for (uint32_t i = threadIdx.x; i < sizeof(ShadeJob) / 4; i += blockDim.x) dst[i] = src[i];
	}
	__syncthreads();
	if (threadIdx.x < 128) sLut[threadIdx.x] = luts[(size_t)sJob.lutIndex * 128 + threadIdx.x];
	__syncthreads();
	const ShadeJob &J = sJob;
	const uint32_t k = (blockIdx.x - J.firstBlock) * 256 + threadIdx.x;
	const uint64_t firstVertex = J.dstVertex + 4ull * k;
	// a lane without a particle (or without room for its quad) stays for the warp-wide stores at the end and shades
	// a record of zeros that is never written
	const bool valid = k < J.numParticles && firstVertex + 4 <= capacityVertices;
	const uint32_t validMask = __ballot_sync(0xffffffffu, valid);
	if (!validMask) return;

	// a_PositionLife / a_VelocitySeed of the quad (its four stream records are identical: read the first)
	const float4 *rec = vertices + 2ull * (J.srcRecord + 4ull * k);
	float4 pl = make_float4(0.0f, 0.0f, 0.0f, 0.0f), vs = pl;
	// one 32-byte record out of every 128-byte line: ask L2 for 64 bytes instead of the whole line
	if (valid) asm("ld.global.L2::64B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(pl.x), "=f"(pl.y), "=f"(pl.z), "=f"(pl.w), "=f"(vs.x), "=f"(vs.y), "=f"(vs.z), "=f"(vs.w) : "l"(rec));
	const float *M = J.instance;        // u_ParticleToWorld, column major
	const float *W = J.instance + 16;   // u_WorldToClip
	const float invDelta = J.instance[36], aspect = J.instance[37];

	// Particle.glsl:78-80
	float px = ((M[0] * pl.x + M[4] * pl.y) + M[8] * pl.z) + M[12];
	float py = ((M[1] * pl.x + M[5] * pl.y) + M[9] * pl.z) + M[13];
	float pz = ((M[2] * pl.x + M[6] * pl.y) + M[10] * pl.z) + M[14];
	float vx = (M[0] * vs.x + M[4] * vs.y) + M[8] * vs.z;
	float vy = (M[1] * vs.x + M[5] * vs.y) + M[9] * vs.z;
	float vz = (M[2] * vs.x + M[6] * vs.y) + M[10] * vs.z;
	const float packedSeed = vs.w;
	// type UBO (std140 offsets of Particle_VertexType_t, src/game/shader/Particle.h:112-125)
	const float *T = J.type;
	const float frameCountX = T[0], frameCountY = T[1], frameRateU = T[2];
	const float rc0 = T[4], rc1 = T[5], rc2 = T[6], rc3 = T[7];
	const float startFrame = T[9];
	const float scaleBase = T[16], scaleVar = T[17], lifeBase = T[18], lifeVar = T[19], relFrameRate = T[20];

	// :82-87
	const float absTotal = lifeBase + lifeVar * packedSeed;
	const float lifeLeft = pl.w + invDelta / absTotal;
	const float lifeTime = 1.0f - lifeLeft;
	const float absLife = lifeTime * absTotal;
	const float s0 = fractf(floorf(packedSeed * 1.0f) * (1.0f / 64.0f));
	const float s1 = fractf(floorf(packedSeed * (1.0f / 64.0f)) * (1.0f / 64.0f));
	const float s2 = fractf(floorf(packedSeed * (1.0f / 4096.0f)) * (1.0f / 64.0f));
	const float s3 = fractf(floorf(packedSeed * (1.0f / 262144.0f)) * (1.0f / 64.0f));
	(void)s1;
	// :89
	px -= vx * invDelta; py -= vy * invDelta; pz -= vz * invDelta;

	// :94-99 LUT rows 0 (spline) and 1 (gradient) at texel coordinate 63 * clamp(lifeTime)
	const float tx = clamp01(lifeTime) * 63.0f;
	const float tf = floorf(tx);
	const uint32_t i0 = min((uint32_t)tf, 63u), i1 = min(i0 + 1u, 63u);
	const float f = tx - tf;
	const uint32_t a0 = sLut[i0], a1 = sLut[i1], g0 = sLut[64 + i0], g1 = sLut[64 + i1];
	const float splX = lut_channel(a0, a1, 0, f), splY = lut_channel(a0, a1, 1, f);
	const float splZ = lut_channel(a0, a1, 2, f), splW = lut_channel(a0, a1, 3, f);
	float cr = srgb_to_linear(lut_channel(g0, g1, 0, f));
	float cg = srgb_to_linear(lut_channel(g0, g1, 1, f));
	float cb = srgb_to_linear(lut_channel(g0, g1, 2, f));

	// :101-107
	float scale = scaleBase + s0 * scaleVar;
	scale *= splX;
	scale = fmaxf(scale, 0.0f);
	const float rotBase = rc0 + rc1 * (s2 * 2.0f - 1.0f);
	const float rotSpin = rc2 + rc3 * (s3 * 2.0f - 1.0f);
	const float rotation = rotBase + rotSpin * absLife;
	const float axx = cosf(rotation), axy = sinf(rotation);
	const float ayx = axy, ayy = -axx;

	// :109
	float nx = ((W[0] * px + W[4] * py) + W[8] * pz) + W[12];
	float ny = ((W[1] * px + W[5] * py) + W[9] * pz) + W[13];
	float nz = ((W[2] * px + W[6] * py) + W[10] * pz) + W[14];
	float nw = ((W[3] * px + W[7] * py) + W[11] * pz) + W[15];

	// :117-125
	const float frameRate = absLife * (1.0f - relFrameRate) + lifeTime * relFrameRate;
	const float frame = frameRate * frameRateU + floorf(s0 * 64.0f + s2 * 4096.0f) * startFrame;
	const float frameI = floorf(frame), frameD = frame - frameI;
	const float f0x = modf_glsl(frameI, frameCountX), f0y = modf_glsl(floorf(frameI / frameCountX), frameCountY);
	const float f1x = modf_glsl(frameI + 1.0f, frameCountX), f1y = modf_glsl(floorf((frameI + 1.0f) / frameCountX), frameCountY);

	// :127 evaluateVisFog (:43-50); no fog image = fully visible (factor (1, 1))
	if (fogTexels) {
		const float *fm = J.instance + 32; // u_VisFogMad
		float2 fac = fog_fetch(fogTexels, fogW, fogH, px * fm[0] + fm[2], pz * fm[1] + fm[3]);
		float t = clamp01((fac.x - 0.45f) * 20.0f);
		float m = fac.y * fac.y * t;
		cr *= m; cg *= m; cb *= m;
	}
	const float erode = splW < 1.0f ? 1.0f / fmaxf(splW, 0.01f) : 1.0f; // :130
	const bool deadQuad = lifeLeft <= 0.0f;                               // :113-115

	// The quad's 256 bytes (4 vertices x {gl_Position, uv, uv | colour, params}) go out through a per-warp staging
	// buffer, two corners at a time: every lane parks its 4 x 32 bytes (144-byte pitch: conflict-free both ways), then
	// store i of lane l writes piece (i * 32 + l) & 3 of the quad of lane (i * 32 + l) >> 2 -- each store instruction
	// of the warp covers 8 complete 128-byte lines instead of one 32-byte piece of 32 different lines
	// (tools/expand_store_bench.cu: the strided form costs 4x the LSU wavefronts for the same bytes).
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	float *stage = sStage[warp];
	float4 *mine = (float4*)(stage + lane * 36);
	float4 *warpDst = out + 4ull * (firstVertex - 4ull * lane); // the quad of lane 0 (lanes hold consecutive particles)
#pragma unroll
	for (uint32_t cp = 0; cp < 2; cp++) {
#pragma unroll
		for (uint32_t c = 0; c < 2; c++) {
			const uint32_t corner = cp * 2 + c;
			const float u = (float)(corner & 1u), v = (float)((corner >> 1) & 1u); // :91
			const float ox = (u * 2.0f - 1.0f) * scale, oy = (v * 2.0f - 1.0f) * scale; // :92, :103
			float4 pos;
			pos.x = nx + (ox * axx + oy * ayx) * (0.5f * aspect); // :111
			pos.y = ny + (ox * axy + oy * ayy) * 0.5f;
			pos.z = nz; pos.w = nw;
			if (deadQuad) pos = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			// per vertex: {gl_Position, v_Uv0, v_Uv1} and {v_Color, v_Params}
			mine[c * 4 + 0] = pos;
			mine[c * 4 + 1] = make_float4((f0x + u) / frameCountX, (f0y + v) / frameCountY, (f1x + u) / frameCountX, (f1y + v) / frameCountY); // :133-134
			mine[c * 4 + 2] = make_float4(cr, cg, cb, splY);      // :129, :135-138
			mine[c * 4 + 3] = make_float4(frameD, splZ, erode, 0.0f);
		}
		__syncwarp();
#pragma unroll
		for (uint32_t i = 0; i < 4; i++) {
			const uint32_t q = i * 32 + lane, prod = q >> 2, piece = q & 3u;
			const float4 *src = (const float4*)(stage + prod * 36 + piece * 8);
			const float4 x = src[0], y = src[1];
			if ((validMask >> prod) & 1u) st256(warpDst + 16ull * prod + cp * 8 + piece * 2, x, y);
		}
		__syncwarp();
	}
}

void launch_shade_vertices(cudaStream_t st, const float4 *vertices, const ShadeJob *jobs, uint32_t numJobs, uint32_t totalBlocks,
	const uint32_t *luts, const uint8_t *fogTexels, uint32_t fogW, uint32_t fogH, float4 *out, uint64_t capacityVertices)
{
	if (!numJobs || !totalBlocks) return;
	k_shade_vertices<<<totalBlocks, 256, 0, st>>>(vertices, jobs, numJobs, luts, fogTexels, fogW, fogH, out, capacityVertices);
}

} // namespace spr
