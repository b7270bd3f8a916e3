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
// Texture filtering follows OpenGL 3.3 section 3.8.11 (LINEAR, the filter of both images: ParticleSystem.cpp:289-290,
// VisFogSystem.cpp:74-75) in float: texel coordinate u * width - 1/2, weights (1-a)(1-b), a(1-b), (1-a)b, ab with
// a, b its fractional parts, UNORM8 texel = c / 255. The spline atlas is sampled at the uv the shader computes from
// u_SplineMad (:403-413); the kernel holds only the type's own two rows, which is all such a uv can reach with a
// non-zero weight (the row coordinate is an exact texel centre). Graphics hardware quantises the weights to 8
// fractional bits; this kernel and its oracles do not. Pinned against the reference's shipped GLSL 3.30 text
// executed on the CPU (oracle/glsl/, tests/test_gpu_vertex_stage.py): bit-identical except where cos / sin / pow
// enter (gl_Position.xy, v_Color.rgb).

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

// The four LINEAR weights of a texel coordinate pair (OpenGL 3.3, 3.8.11) and the weighted sum in the order the
// specification writes it: tau = (1-a)(1-b) t00 + a(1-b) t10 + (1-a)b t01 + ab t11.
struct BilinearWeights { float w00, w10, w01, w11; };
__device__ __forceinline__ BilinearWeights bilinear_weights(float a, float b)
{
	BilinearWeights w;
	w.w00 = (1.0f - a) * (1.0f - b); w.w10 = a * (1.0f - b); w.w01 = (1.0f - a) * b; w.w11 = a * b;
	return w;
}
__device__ __forceinline__ float bilinear(const BilinearWeights &w, float t00, float t10, float t01, float t11)
{
	return ((t00 * w.w00 + t10 * w.w10) + t01 * w.w01) + t11 * w.w11;
}
// UNORM8 -> float: v / 255 correctly rounded, without the division: q = v * (1/255) is off by at most one ulp, one
// Newton step on the remainder (two FMAs) lands on the IEEE quotient -- equal to v / 255.0f for all 256 values of v
// (checked exhaustively on the host: tests/test_abi.py::test_unorm8_division_free).
__device__ __forceinline__ float unorm8(uint32_t v)
{
	const float r = 1.0f / 255.0f;
	const float x = (float)v, q = x * r;
	return __fmaf_rn(__fmaf_rn(-q, 255.0f, x), r, q);
}
__device__ __forceinline__ float texel_channel(uint32_t t, int c) { return unorm8((t >> (8 * c)) & 255u); }

// LINEAR fetch of an RG8 image with clamp-to-edge addressing; returns (r, g)
__device__ inline float2 fog_fetch(const uint8_t *texels, uint32_t w, uint32_t h, float u, float v)
{
	float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
	float fx0 = floorf(x), fy0 = floorf(y);
	const BilinearWeights bw = bilinear_weights(x - fx0, y - fy0);
	int x0 = (int)fx0, y0 = (int)fy0;
	int x1 = x0 + 1, y1 = y0 + 1;
	x0 = min(max(x0, 0), (int)w - 1); x1 = min(max(x1, 0), (int)w - 1);
	y0 = min(max(y0, 0), (int)h - 1); y1 = min(max(y1, 0), (int)h - 1);
	float c[2];
#pragma unroll
	for (int k = 0; k < 2; k++) {
		float t00 = unorm8(texels[2 * ((size_t)y0 * w + x0) + k]);
		float t10 = unorm8(texels[2 * ((size_t)y0 * w + x1) + k]);
		float t01 = unorm8(texels[2 * ((size_t)y1 * w + x0) + k]);
		float t11 = unorm8(texels[2 * ((size_t)y1 * w + x1) + k]);
		c[k] = bilinear(bw, t00, t10, t01, t11);
	}
	return make_float2(c[0], c[1]);
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
	const float splineMadX = T[12], splineMadY = T[13], splineMadZ = T[14], splineMadW = T[15];
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

	// :94-99 texture(u_SplineTexture, splineUv) and the gradient one row below, LINEAR over the 512 x 512 atlas
	// (ParticleSystem.cpp:253-258). The type's cell starts at the texel whose centre is u_SplineMad.zw (:406-413);
	// rows and columns are taken relative to it and clamped to the cell -- a neighbour outside it only ever has
	// weight 0 (column 64 at lifeTime = 1, the row under the gradient).
	const float kAtlasW = (float)(kSplineSamples * kSplineAtlasTypesX), kAtlasH = (float)(kSplineAtlasTypesY * kSplineRowsPerType);
	const int cellX = (int)floorf(splineMadZ * kAtlasW), cellY = (int)floorf(splineMadW * kAtlasH);
	const float su = clamp01(lifeTime) * splineMadX + splineMadZ;
	const float sv = 0.0f * splineMadY + splineMadW;
	const float gv = sv + splineMadY;                                 // splineUv + vec2(0, u_SplineMad.y)
	const float sx = su * kAtlasW - 0.5f, sy = sv * kAtlasH - 0.5f, gy = gv * kAtlasH - 0.5f;
	const float sfx = floorf(sx), sfy = floorf(sy), gfy = floorf(gy);
	const int c0 = min(max((int)sfx - cellX, 0), 63), c1 = min(max((int)sfx + 1 - cellX, 0), 63);
	const int sr0 = min(max((int)sfy - cellY, 0), 1), sr1 = min(max((int)sfy + 1 - cellY, 0), 1);
	const int gr0 = min(max((int)gfy - cellY, 0), 1), gr1 = min(max((int)gfy + 1 - cellY, 0), 1);
	const float ca = sx - sfx, sb = sy - sfy, gb = gy - gfy;
	float splX, splY, splZ, splW, cr, cg, cb;
	if (sb == 0.0f && gb == 0.0f) {
		// The row coordinate is an exact texel centre (always, for a UBO built as :406-413 builds it; uniform over the
		// grid's job): the lower row's weights are (1-a) * 0 and a * 0, its terms add +0 to a non-negative sum, and the
		// upper row's weights are (1-a) * 1 and a * 1. Same bits as the four-texel sum below with half the fetches.
		const float w0 = 1.0f - ca, w1 = ca;
		const uint32_t s0t = sLut[sr0 * 64 + c0], s1t = sLut[sr0 * 64 + c1], g0t = sLut[gr0 * 64 + c0], g1t = sLut[gr0 * 64 + c1];
#define SPR_SAMPLE2(t0, t1, c) (texel_channel(t0, c) * w0 + texel_channel(t1, c) * w1)
		splX = SPR_SAMPLE2(s0t, s1t, 0); splY = SPR_SAMPLE2(s0t, s1t, 1); splZ = SPR_SAMPLE2(s0t, s1t, 2); splW = SPR_SAMPLE2(s0t, s1t, 3);
		cr = SPR_SAMPLE2(g0t, g1t, 0); cg = SPR_SAMPLE2(g0t, g1t, 1); cb = SPR_SAMPLE2(g0t, g1t, 2);
#undef SPR_SAMPLE2
	} else {
		const BilinearWeights sw = bilinear_weights(ca, sb), gw = bilinear_weights(ca, gb);
		const uint32_t s00 = sLut[sr0 * 64 + c0], s10 = sLut[sr0 * 64 + c1], s01 = sLut[sr1 * 64 + c0], s11 = sLut[sr1 * 64 + c1];
		const uint32_t g00 = sLut[gr0 * 64 + c0], g10 = sLut[gr0 * 64 + c1], g01 = sLut[gr1 * 64 + c0], g11 = sLut[gr1 * 64 + c1];
#define SPR_SAMPLE(w, t00, t10, t01, t11, c) bilinear(w, texel_channel(t00, c), texel_channel(t10, c), texel_channel(t01, c), texel_channel(t11, c))
		splX = SPR_SAMPLE(sw, s00, s10, s01, s11, 0); splY = SPR_SAMPLE(sw, s00, s10, s01, s11, 1);
		splZ = SPR_SAMPLE(sw, s00, s10, s01, s11, 2); splW = SPR_SAMPLE(sw, s00, s10, s01, s11, 3);
		cr = SPR_SAMPLE(gw, g00, g10, g01, g11, 0); cg = SPR_SAMPLE(gw, g00, g10, g01, g11, 1); cb = SPR_SAMPLE(gw, g00, g10, g01, g11, 2);
#undef SPR_SAMPLE
	}
	cr = srgb_to_linear(cr); cg = srgb_to_linear(cg); cb = srgb_to_linear(cb);

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
