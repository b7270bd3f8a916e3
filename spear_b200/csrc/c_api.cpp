// c_api.cpp -- the extern "C" boundary declared in include/spear_particles.h.
// Thin: argument checks, exception -> status translation, read-back copies.

#include "particle_system.h"

#include <string.h>

#include <new>

using namespace spr;

struct SprBillboardSystem { BillboardSystem sys; SprBillboardSystem(Context *c, uint32_t n) : sys(c, n) { } };
struct SprContext { Context ctx; explicit SprContext(const SprContextDesc &d) : ctx(d) { } };
struct SprSystem { ParticleSystem sys; SprSystem(Context *c, const SprSystemsDesc &d) : sys(c, d) { } };

static thread_local std::string g_lastError;

template <typename F>
static int guarded(F &&f)
{
	try { f(); return SPR_OK; }
	catch (const CudaError &e) { g_lastError = e.what(); return SPR_ERR_CUDA; }
	catch (const Unsupported &e) { g_lastError = e.what(); return SPR_ERR_UNSUPPORTED; }
	catch (const std::bad_alloc &) { g_lastError = "out of host memory"; return SPR_ERR_OUT_OF_MEMORY; }
	catch (const std::exception &e) { g_lastError = e.what(); return SPR_ERR_INVALID_ARGUMENT; }
}

#define REQUIRE(cond) do { if (!(cond)) { g_lastError = "invalid argument: " #cond; return SPR_ERR_INVALID_ARGUMENT; } } while (0)

static bool validEffect(SprSystem *s, uint32_t id) { return s && id < s->sys.effects.size() && s->sys.effects[id].inUse; }

extern "C" {

SPR_API void sprComponentDefaults(SprParticleSystemComponent *c)
{
	memset(c, 0, sizeof(*c));
	c->frameCount[0] = c->frameCount[1] = 1;
	c->timeStep = 0.1f;
	c->updateRadius = 10.0f;
	c->cullPadding = 1.5f;
	c->spawnTime = 0.2f;
	c->emitterOnTime = -1.0f;
	SprRandomVec3 *rv[2] = { &c->emitPosition, &c->emitVelocity };
	for (int i = 0; i < 2; i++) {
		rv[i]->sphere.maxTheta = 360.0f;
		rv[i]->sphere.maxPhi = 180.0f;
		rv[i]->sphere.scale.x = rv[i]->sphere.scale.y = rv[i]->sphere.scale.z = 1.0f;
	}
	c->size = 0.5f;
	c->lifeTime = 1.0f;
	c->gradient.defaultColor.x = c->gradient.defaultColor.y = c->gradient.defaultColor.z = 1.0f;
}

SPR_API const char *sprGetLastError(void) { return g_lastError.c_str(); }

SPR_API int sprContextCreate(const SprContextDesc *desc, SprContext **out)
{
	REQUIRE(desc && out);
	*out = nullptr;
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { g_lastError = "no CUDA device available (this library has no CPU fallback)"; return SPR_ERR_NO_DEVICE; }
	REQUIRE(desc->device >= 0 && desc->device < count);
	return guarded([&] { *out = new SprContext(*desc); });
}

SPR_API int sprContextDestroy(SprContext *ctx) { REQUIRE(ctx); return guarded([&] { delete ctx; }); }
SPR_API int sprContextSynchronize(SprContext *ctx) { REQUIRE(ctx); return guarded([&] { ctx->ctx.sync(); }); }
SPR_API void *sprContextStream(SprContext *ctx) { return ctx ? (void*)ctx->ctx.stream : nullptr; }
SPR_API uint64_t sprContextLaunchCount(SprContext *ctx) { return ctx ? ctx->ctx.launchCount : 0; }

static const char *const kKernelNames[17] = {
	"k_scatter_words", "k_plan", "k_emit_warp", "k_emit_burst", "k_integrate", "k_finalize", "k_stream_fused", "k_gather_render",
	"k_pack_runs", "k_expand_runs", "k_sort_hist", "k_sort_scan", "k_sort_pass", "k_expand_sorted", "k_dead_append", "k_expand_stream", "k_sort_small" };

SPR_API int sprContextEnableKernelTiming(SprContext *ctx, int enable)
{
	REQUIRE(ctx);
	return guarded([&] { ctx->ctx.resolveTimings(); ctx->ctx.timingEnabled = enable != 0; });
}

SPR_API int sprContextGetKernelTimes(SprContext *ctx, int reset, const char **names, double *ms, uint64_t *launches, uint32_t capacity, uint32_t *outCount)
{
	REQUIRE(ctx && outCount);
	return guarded([&] {
		ctx->ctx.resolveTimings();
		uint32_t n = 0;
		for (int k = 0; k < 17; k++) {
			if (!ctx->ctx.kernelLaunches[k]) continue;
			if (n < capacity) { if (names) names[n] = kKernelNames[k]; if (ms) ms[n] = ctx->ctx.kernelMs[k]; if (launches) launches[n] = ctx->ctx.kernelLaunches[k]; }
			n++;
		}
		*outCount = n;
		if (reset) for (int k = 0; k < 17; k++) { ctx->ctx.kernelMs[k] = 0; ctx->ctx.kernelLaunches[k] = 0; }
	});
}

SPR_API int sprSystemCreate(SprContext *ctx, const SprSystemsDesc *desc, SprSystem **out)
{
	REQUIRE(ctx && desc && out);
	return guarded([&] { *out = new SprSystem(&ctx->ctx, *desc); });
}

SPR_API int sprContextGetTotals(SprContext *ctx, uint64_t *outSimSteps, uint64_t *outAliveParticles)
{
	REQUIRE(ctx);
	return guarded([&] {
		if (outSimSteps) *outSimSteps = ctx->ctx.totalSimSteps;
		if (outAliveParticles) {
			std::vector<uint32_t> globals;
			for (uint32_t g = 0; g < ctx->ctx.numEffectGlobals; g++) if (ctx->ctx.effectLive[g]) globals.push_back(g);
			std::vector<RenderGather> got(globals.size());
			ctx->ctx.gatherEffects(globals.data(), (uint32_t)globals.size(), got.data());
			uint64_t total = 0;
			for (const RenderGather &r : got) total += r.aliveCount;
			*outAliveParticles = total;
		}
	});
}

SPR_API int sprSystemDestroy(SprSystem *sys) { REQUIRE(sys); return guarded([&] { delete sys; }); }

SPR_API int sprReserveEffectType(SprSystem *sys, const SprParticleSystemComponent *c, uint32_t *outTypeId)
{
	REQUIRE(sys && c);
	return guarded([&] { uint32_t t = sys->sys.reserveEffectType(c); if (outTypeId) *outTypeId = t; });
}

SPR_API int sprReleaseEffectType(SprSystem *sys, const SprParticleSystemComponent *c)
{
	REQUIRE(sys && c);
	return guarded([&] { sys->sys.releaseEffectType(c); });
}

SPR_API int sprAddEffect(SprSystem *sys, uint32_t entityId, uint8_t componentIndex,
	const SprParticleSystemComponent *c, const SprTransform *transform, uint32_t *outEffectId)
{
	REQUIRE(sys && c && transform);
	return guarded([&] { uint32_t e = sys->sys.addEffect(entityId, componentIndex, c, *transform); if (outEffectId) *outEffectId = e; });
}

SPR_API int sprUpdateTransform(SprSystem *sys, uint32_t effectId, const SprTransformUpdate *update)
{
	REQUIRE(validEffect(sys, effectId) && update);
	return guarded([&] { sys->sys.updateTransform(effectId, *update); });
}

SPR_API int sprPrepareForRemove(SprSystem *sys, uint32_t effectId, const SprFrameArgs *args, int *outCanRemove)
{
	REQUIRE(validEffect(sys, effectId) && args);
	return guarded([&] { bool r = sys->sys.prepareForRemove(effectId, *args); if (outCanRemove) *outCanRemove = r ? 1 : 0; });
}

SPR_API int sprRemoveEffect(SprSystem *sys, uint32_t effectId)
{
	REQUIRE(validEffect(sys, effectId));
	return guarded([&] { sys->sys.remove(effectId); });
}

SPR_API int sprUpdateParticles(SprSystem *sys, const uint32_t *activeIds, uint32_t numActive, const SprFrameArgs *args)
{
	REQUIRE(sys && args && (activeIds || numActive == 0));
	return guarded([&] { sys->sys.updateParticles(activeIds, numActive, *args); }); // ids are checked by Context::validateRequests
}

SPR_API int sprUpdateParticlesBatch(SprContext *ctx, SprSystem *const *systems, uint32_t numSystems,
	const uint32_t *const *activeIds, const uint32_t *numActive, const SprFrameArgs *args, size_t argsStride)
{
	REQUIRE(ctx && systems && activeIds && numActive && args);
	return guarded([&] {
		std::vector<UpdateRequest> reqs(numSystems);
		for (uint32_t i = 0; i < numSystems; i++) {
			if (!systems[i] || systems[i]->sys.ctx != &ctx->ctx) throw std::runtime_error("system does not belong to this context");
			reqs[i].system = &systems[i]->sys;
			reqs[i].activeIds = activeIds[i];
			reqs[i].numActive = numActive[i];
			reqs[i].args = (const SprFrameArgs*)((const char*)args + argsStride * i);
		}
		ctx->ctx.update(reqs.data(), numSystems);
	});
}

SPR_API int sprRenderMain(SprSystem *sys, const uint32_t *visibleIds, uint32_t numVisible, const SprRenderArgs *args, SprRenderOutput *out)
{
	REQUIRE(sys && args && out && (visibleIds || numVisible == 0));
	for (uint32_t i = 0; i < numVisible; i++) REQUIRE(validEffect(sys, visibleIds[i]));
	return guarded([&] { sys->sys.renderMain(visibleIds, numVisible, *args, *out); });
}

SPR_API int sprRenderMainBatch(SprContext *ctx, SprSystem *const *systems, uint32_t numSystems,
	const uint32_t *const *visibleIds, const uint32_t *numVisible, const SprRenderArgs *args, size_t argsStride, SprRenderOutput *outs)
{
	REQUIRE(ctx && systems && visibleIds && numVisible && args && outs);
	return guarded([&] {
		std::vector<ParticleSystem*> sys(numSystems);
		for (uint32_t i = 0; i < numSystems; i++) {
			if (!systems[i] || systems[i]->sys.ctx != &ctx->ctx) throw std::runtime_error("system does not belong to this context");
			for (uint32_t k = 0; k < numVisible[i]; k++) if (!validEffect(systems[i], visibleIds[i][k])) throw std::runtime_error("invalid effect id");
			sys[i] = &systems[i]->sys;
		}
		ctx->ctx.renderBatch(sys.data(), numSystems, visibleIds, numVisible, args, argsStride, outs);
	});
}

// ---- parity read-backs

SPR_API int sprGetEffectInfo(SprSystem *sys, uint32_t effectId, SprEffectInfo *out)
{
	REQUIRE(validEffect(sys, effectId) && out);
	return guarded([&] {
		const HostEffect &E = sys->sys.effects[effectId];
		RenderGather g;
		sys->sys.ctx->gatherEffects(&E.global, 1, &g);
		memset(out, 0, sizeof(*out));
		out->typeId = E.typeId;
		out->slotCount = g.slotCount;
		out->aliveCount = g.aliveCount;
		out->freeCount = g.freeCount;
		out->timeStep = E.timeStep;
		out->timeDelta = E.timeDelta;
		out->spawnTimer = g.spawnTimer;
		out->emitOnTimer = g.emitOnTimer;
		out->stopEmit = (g.stopEmit || E.hostStopEmit) ? 1 : 0;
		out->firstEmit = (uint8_t)g.firstEmit;
		out->instantDelete = E.instantDelete ? 1 : 0;
		out->gpuBounds = boundsFromGather(g, sys->sys.types[E.typeId], sys->sys.effectsCold[effectId]);
		out->lastUpdateTime = E.lastUpdateTime;
		out->simSteps = E.simSteps;
	});
}

static int readRange(SprSystem *sys, uint32_t effectId, int what, void *out, uint32_t capacity, uint32_t *outCount)
{
	return guarded([&] {
		Context *ctx = sys->sys.ctx;
		SPR_CUDA(cudaSetDevice(ctx->device));
		const HostEffect &E = sys->sys.effects[effectId];
		RenderGather g;
		ctx->gatherEffects(&E.global, 1, &g);
		const EffectParams &P = ctx->hParams[E.global];
		uint32_t n = 0; size_t elem = 0; const void *src = nullptr;
		switch (what) {
		case 0: n = g.freeCount; elem = 4; src = ctx->dFreeStack + P.slotBase; break;
		case 1: n = g.aliveCount * 4; elem = 32; src = ctx->dVertices + 8ull * P.slotBase; break;
		case 2: n = g.slotCount; elem = 32; src = ctx->dSlotState + 2ull * P.slotBase; break;
		}
		if (outCount) *outCount = n;
		uint32_t m = n < capacity ? n : capacity;
		if (m && out) {
			SPR_CUDA(cudaMemcpyAsync(out, src, (size_t)m * elem, cudaMemcpyDeviceToHost, ctx->stream));
			SPR_CUDA(cudaStreamSynchronize(ctx->stream));
		}
	});
}

SPR_API int sprReadFreeIndices(SprSystem *sys, uint32_t effectId, uint32_t *out, uint32_t capacity, uint32_t *outCount)
{
	REQUIRE(validEffect(sys, effectId));
	return readRange(sys, effectId, 0, out, capacity, outCount);
}

SPR_API int sprReadVertexStream(SprSystem *sys, uint32_t effectId, SprGpuParticle *out, uint32_t capacityVertices, uint32_t *outCount)
{
	REQUIRE(validEffect(sys, effectId));
	return readRange(sys, effectId, 1, out, capacityVertices, outCount);
}

SPR_API int sprReadSlots(SprSystem *sys, uint32_t effectId, SprGpuParticle *out, uint32_t capacitySlots, uint32_t *outCount)
{
	REQUIRE(validEffect(sys, effectId));
	return readRange(sys, effectId, 2, out, capacitySlots, outCount);
}

SPR_API int sprReadStreamSlots(SprSystem *sys, uint32_t effectId, uint32_t *out, uint32_t capacity, uint32_t *outCount)
{
	REQUIRE(validEffect(sys, effectId));
	return guarded([&] {
		Context *ctx = sys->sys.ctx;
		const HostEffect &E = sys->sys.effects[effectId];
		RenderGather g;
		ctx->gatherEffects(&E.global, 1, &g);
		const EffectParams &P = ctx->hParams[E.global];
		if (outCount) *outCount = g.aliveCount;
		uint32_t m = g.aliveCount < capacity ? g.aliveCount : capacity;
		if (!m || !out) return;
		if (ctx->flags & SPR_CTX_SORT_PARTICLES) {
			SPR_CUDA(cudaMemcpyAsync(out, ctx->dVals[0] + P.slotBase, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
			SPR_CUDA(cudaStreamSynchronize(ctx->stream));
		} else {
			// stream order is ascending slot order over the live slots (ParticleSystem.cpp:611-665)
			std::vector<SprGpuParticle> slots(g.slotCount);
			if (g.slotCount) {
				SPR_CUDA(cudaMemcpyAsync(slots.data(), ctx->dSlotState + 2ull * P.slotBase, (size_t)g.slotCount * 32, cudaMemcpyDeviceToHost, ctx->stream));
				SPR_CUDA(cudaStreamSynchronize(ctx->stream));
			}
			uint32_t k = 0;
			for (uint32_t i = 0; i < g.slotCount && k < m; i++) if (slots[i].life < kHugeLifeCmp) out[k++] = i;
		}
	});
}

SPR_API int sprGetRngState(SprSystem *sys, uint64_t outStateInc[2])
{
	REQUIRE(sys && outStateInc);
	return guarded([&] {
		Context *ctx = sys->sys.ctx;
		ctx->flushUploads();
		SystemDev sd;
		SPR_CUDA(cudaMemcpyAsync(&sd, ctx->dSystems.ptr + sys->sys.systemIdx, sizeof(sd), cudaMemcpyDeviceToHost, ctx->stream));
		SPR_CUDA(cudaStreamSynchronize(ctx->stream));
		outStateInc[0] = sd.rngState; outStateInc[1] = sd.rngInc;
	});
}

SPR_API int sprDescriptorsFromJson(const char *json, size_t length, SprDescriptorSet **out)
{
	REQUIRE(json && out);
	return guarded([&] {
		std::string error;
		DescriptorSet *set = descriptorsFromJson(json, length, error);
		if (!set) throw std::invalid_argument(error);
		*out = (SprDescriptorSet*)set;
	});
}

SPR_API uint32_t sprDescriptorCount(const SprDescriptorSet *set) { return set ? descriptorCount((const DescriptorSet*)set) : 0; }
SPR_API const SprParticleSystemComponent *sprDescriptorGet(const SprDescriptorSet *set, uint32_t index) { return set ? descriptorGet((const DescriptorSet*)set, index) : nullptr; }
SPR_API const char *sprDescriptorTextureName(const SprDescriptorSet *set, uint32_t index) { return set ? descriptorTexture((const DescriptorSet*)set, index) : nullptr; }
SPR_API void sprDescriptorsFree(SprDescriptorSet *set) { if (set) descriptorsFree((DescriptorSet*)set); }

SPR_API int sprQueryEffects(SprSystem *sys, const SprFrustum *frustum, uint32_t *outIds, uint32_t capacity, uint32_t *outCount)
{
	REQUIRE(sys && frustum && (outIds || !capacity));
	return guarded([&] {
		uint32_t n = sys->sys.queryEffects(*frustum, outIds, capacity);
		if (outCount) *outCount = n;
	});
}

SPR_API int sprShadeVertices(SprSystem *sys, const SprDrawRecord *records, uint32_t numRecords, const SprFogImage *fog,
	SprShadedVertex *deviceOut, uint64_t capacityVertices, uint64_t *outNumVertices)
{
	REQUIRE(sys && (records || !numRecords) && (deviceOut || !capacityVertices));
	REQUIRE(!fog || (fog->deviceTexels && fog->width && fog->height));
	return guarded([&] {
		uint64_t n = sys->sys.shadeVertices(records, numRecords, fog, deviceOut, capacityVertices);
		if (outNumVertices) *outNumVertices = n;
	});
}

SPR_API int sprBillboardSystemCreate(SprContext *ctx, uint32_t maxBillboardsPerFrame, SprBillboardSystem **out)
{
	REQUIRE(ctx && out);
	return guarded([&] { *out = new SprBillboardSystem(&ctx->ctx, maxBillboardsPerFrame); });
}

SPR_API int sprBillboardSystemDestroy(SprBillboardSystem *sys)
{
	REQUIRE(sys);
	return guarded([&] { cudaSetDevice(sys->sys.ctx->device); cudaStreamSynchronize(sys->sys.ctx->stream); delete sys; });
}

SPR_API int sprBillboardAdd(SprBillboardSystem *sys, const SprBillboard *billboards, uint32_t count)
{
	REQUIRE(sys && (billboards || !count));
	return guarded([&] { sys->sys.add(billboards, count); });
}

SPR_API int sprBillboardPack(const SprBillboard *billboards, uint32_t count, SprBillboardPacked *out, uint32_t *outCount)
{
	REQUIRE((billboards && out) || !count);
	return guarded([&] {
		BillboardDev *end = BillboardSystem::pack(billboards, 0, count, (BillboardDev*)out);
		if (outCount) *outCount = (uint32_t)(end - (BillboardDev*)out);
	});
}

SPR_API int sprBillboardAddPacked(SprBillboardSystem *sys, const SprBillboardPacked *packed, uint32_t count, int memory)
{
	REQUIRE(sys && (packed || !count) && (memory == SPR_MEMORY_HOST || memory == SPR_MEMORY_DEVICE));
	return guarded([&] { sys->sys.addPacked((const BillboardDev*)packed, count, memory == SPR_MEMORY_DEVICE); });
}

SPR_API int sprBillboardRenderMain(SprBillboardSystem *sys, const SprRenderArgs *args, SprBillboardOutput *out)
{
	REQUIRE(sys && args && out);
	return guarded([&] { sys->sys.renderMain(*args, *out); });
}

SPR_API int sprBillboardReadVertices(SprBillboardSystem *sys, SprBillboardVertex *out, uint32_t capacityQuads, uint32_t *outQuads)
{
	REQUIRE(sys && (out || !capacityQuads));
	return guarded([&] {
		BillboardSystem &b = sys->sys;
		uint32_t n = b.lastQuads;
		if (outQuads) *outQuads = n;
		uint32_t m = n < capacityQuads ? n : capacityQuads;
		if (m) {
			SPR_CUDA(cudaSetDevice(b.ctx->device));
			SPR_CUDA(cudaMemcpyAsync(out, b.dVerts.ptr, (size_t)m * 4 * sizeof(SprBillboardVertex), cudaMemcpyDeviceToHost, b.ctx->stream));
			SPR_CUDA(cudaStreamSynchronize(b.ctx->stream));
		}
	});
}

SPR_API int sprGetEffectTypeLut(SprSystem *sys, uint32_t typeId, uint8_t out[SPR_SPLINE_LUT_BYTES])
{
	REQUIRE(sys && out && typeId < sys->sys.types.size() && sys->sys.types[typeId].comp);
	memcpy(out, sys->sys.types[typeId].lut, SPR_SPLINE_LUT_BYTES);
	return SPR_OK;
}

SPR_API int sprGetEffectTypeUbo(SprSystem *sys, uint32_t typeId, SprVertexTypeUbo *out)
{
	REQUIRE(sys && out && typeId < sys->sys.types.size() && sys->sys.types[typeId].comp);
	*out = sys->sys.types[typeId].typeUbo;
	return SPR_OK;
}

SPR_API int sprBakeEffectType(const SprParticleSystemComponent *c, uint32_t typeId, uint8_t outLut[SPR_SPLINE_LUT_BYTES], SprVertexTypeUbo *outUbo)
{
	REQUIRE(c && outLut && outUbo);
	bakeTypeUbo(*outUbo, typeId, *c);
	memset(outLut, 0, SPR_SPLINE_LUT_BYTES);
	bakeSplineRow(outLut, 0, c->scaleSpline, false);
	bakeSplineRow(outLut, 1, c->alphaSpline, false);
	bakeSplineRow(outLut, 2, c->additiveSpline, true);
	bakeSplineRow(outLut, 3, c->erosionSpline, false);
	bakeGradientRow(outLut + SPR_SPLINE_SAMPLE_RATE * 4, c->gradient);
	return SPR_OK;
}

SPR_API int sprGetEffectDeviceView(SprSystem *sys, uint32_t effectId, SprEffectDeviceView *out)
{
	REQUIRE(validEffect(sys, effectId) && out);
	Context *ctx = sys->sys.ctx;
	const HostEffect &E = sys->sys.effects[effectId];
	const EffectParams &P = ctx->hParams[E.global];
	out->vertices = (const SprGpuParticle*)(ctx->dVertices + 8ull * P.slotBase);
	out->slots = (const SprGpuParticle*)(ctx->dSlotState + 2ull * P.slotBase);
	out->aliveCount = &ctx->dState.ptr[E.global].aliveCount;
	out->vertexByteOffset = (uint64_t)P.slotBase * 128ull;
	out->slotCapacity = P.slotCapacity;
	return SPR_OK;
}

SPR_API int sprPackRuns(SprContext *c, SprSystem *const *systems, const uint32_t *effectIds, uint32_t numEffects,
	SprGpuParticle *deviceOut, uint64_t capacityRecords, uint32_t *deviceCounts, uint64_t *outTotalRecords)
{
	REQUIRE(c && systems && effectIds && deviceOut);
	return guarded([&] {
		Context *ctx = &c->ctx;
		std::vector<uint32_t> globals(numEffects);
		for (uint32_t i = 0; i < numEffects; i++) {
			if (!validEffect(systems[i], effectIds[i])) throw std::runtime_error("invalid effect in sprPackRuns");
			globals[i] = systems[i]->sys.effects[effectIds[i]].global;
		}
		ctx->flushUploads();
		ctx->gatherIdxD.reserve(numEffects, 0, ctx->stream);
		ctx->dRunOffsets.reserve(numEffects + 1, 0, ctx->stream);
		SPR_CUDA(cudaMemcpyAsync(ctx->gatherIdxD.ptr, globals.data(), numEffects * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
		launch_run_offsets(ctx->stream, ctx->tablePtrs(), ctx->gatherIdxD.ptr, numEffects, ctx->dRunOffsets.ptr, deviceCounts);
		launch_pack_runs(ctx->stream, ctx->poolPtrs(), ctx->tablePtrs(), ctx->gatherIdxD.ptr, numEffects, ctx->dRunOffsets.ptr,
			(float4*)deviceOut, capacityRecords, ctx->numSms, ctx->deferExpansion);
		ctx->launchCount += 2;
		SPR_CUDA(cudaGetLastError());
		if (outTotalRecords) {
			SPR_CUDA(cudaMemcpyAsync(outTotalRecords, ctx->dRunOffsets.ptr + numEffects, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
			SPR_CUDA(cudaStreamSynchronize(ctx->stream));
		}
	});
}

struct SprRunBuffer { Context *ctx; void *dev; uint64_t capacityRecords; };

SPR_API int sprRunBufferCreate(SprContext *c, uint64_t capacityRecords, SprRunBuffer **out)
{
	REQUIRE(c && out && capacityRecords);
	return guarded([&] {
		SPR_CUDA(cudaSetDevice(c->ctx.device));
		SprRunBuffer *b = new SprRunBuffer{ &c->ctx, nullptr, capacityRecords };
		SPR_CUDA(cudaMalloc(&b->dev, 32 + capacityRecords * 32)); // plain cudaMalloc: exportable through CUDA IPC
		SPR_CUDA(cudaMemsetAsync(b->dev, 0, 32, c->ctx.stream));
		*out = b;
	});
}

SPR_API int sprRunBufferDestroy(SprRunBuffer *buf)
{
	REQUIRE(buf);
	return guarded([&] { SPR_CUDA(cudaSetDevice(buf->ctx->device)); cudaStreamSynchronize(buf->ctx->stream); cudaFree(buf->dev); delete buf; });
}

SPR_API void *sprRunBufferDevicePtr(SprRunBuffer *buf) { return buf ? buf->dev : nullptr; }

SPR_API int sprRunBufferGetIpcHandle(SprRunBuffer *buf, uint8_t outHandle[SPR_IPC_HANDLE_BYTES])
{
	REQUIRE(buf && outHandle);
	static_assert(sizeof(cudaIpcMemHandle_t) == SPR_IPC_HANDLE_BYTES, "ipc handle size");
	return guarded([&] { SPR_CUDA(cudaSetDevice(buf->ctx->device)); cudaIpcMemHandle_t h; SPR_CUDA(cudaIpcGetMemHandle(&h, buf->dev)); memcpy(outHandle, &h, sizeof(h)); });
}

SPR_API int sprRunBufferOpenPeer(SprContext *c, const uint8_t handle[SPR_IPC_HANDLE_BYTES], void **outDevicePtr)
{
	REQUIRE(c && handle && outDevicePtr);
	return guarded([&] {
		SPR_CUDA(cudaSetDevice(c->ctx.device));
		cudaIpcMemHandle_t h; memcpy(&h, handle, sizeof(h));
		SPR_CUDA(cudaIpcOpenMemHandle(outDevicePtr, h, cudaIpcMemLazyEnablePeerAccess));
	});
}

SPR_API int sprRunBufferClosePeer(SprContext *c, void *devicePtr)
{
	REQUIRE(c && devicePtr);
	return guarded([&] { SPR_CUDA(cudaSetDevice(c->ctx.device)); cudaStreamSynchronize(c->ctx.stream); SPR_CUDA(cudaIpcCloseMemHandle(devicePtr)); });
}

SPR_API int sprRunBufferPack(SprContext *c, SprSystem *const *systems, const uint32_t *effectIds, uint32_t numEffects, SprRunBuffer *buf)
{
	REQUIRE(c && systems && effectIds && buf);
	return guarded([&] {
		Context *ctx = &c->ctx;
		// the effect list of a steady-state frame does not change: keep the uploaded list of the last call
		std::vector<uint32_t> globals(numEffects);
		for (uint32_t i = 0; i < numEffects; i++) {
			if (!validEffect(systems[i], effectIds[i])) throw std::runtime_error("invalid effect in sprRunBufferPack");
			globals[i] = systems[i]->sys.effects[effectIds[i]].global;
		}
		ctx->flushUploads();
		if (globals != ctx->packList) {
			ctx->packListD.reserve(numEffects, 0, ctx->stream);
			ctx->dRunOffsets.reserve(numEffects + 1, 0, ctx->stream);
			SPR_CUDA(cudaMemcpyAsync(ctx->packListD.ptr, globals.data(), numEffects * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
			SPR_CUDA(cudaStreamSynchronize(ctx->stream));
			ctx->packList.swap(globals);
		}
		launch_run_offsets(ctx->stream, ctx->tablePtrs(), ctx->packListD.ptr, numEffects, ctx->dRunOffsets.ptr, nullptr);
		launch_pack_runs_buf(ctx->stream, ctx->poolPtrs(), ctx->tablePtrs(), ctx->packListD.ptr, numEffects, ctx->dRunOffsets.ptr,
			buf->dev, buf->capacityRecords, ctx->numSms, ctx->deferExpansion);
		ctx->launchCount += 2;
		SPR_CUDA(cudaGetLastError());
	});
}

SPR_API int sprExpandFromPeers(SprContext *c, void *const *bufferPtrs, uint32_t numRanks, uint32_t myRank, SprGpuParticle *deviceVerticesOut,
	uint64_t capacityRecords, uint64_t *deviceTotalRecords)
{
	REQUIRE(c && bufferPtrs && deviceVerticesOut && numRanks >= 1 && numRanks <= 16 && myRank < numRanks);
	return guarded([&] {
		SPR_CUDA(cudaSetDevice(c->ctx.device));
		launch_expand_from_peers(c->ctx.stream, bufferPtrs, numRanks, myRank, (float4*)deviceVerticesOut, capacityRecords, deviceTotalRecords, c->ctx.numSms * c->ctx.assemblyCtasPerSm,
			c->ctx.assemblyKernel ? c->ctx.assemblyKernel == 2 : c->ctx.assemblyCtasPerSm <= 2);
		c->ctx.launchCount++;
		SPR_CUDA(cudaGetLastError());
	});
}

SPR_API int sprContextSetDeferredExpansion(SprContext *c, int on)
{
	REQUIRE(c);
	if (on && !(c->ctx.flags & SPR_CTX_SORT_PARTICLES)) { g_lastError = "deferred expansion needs SPR_CTX_SORT_PARTICLES"; return SPR_ERR_UNSUPPORTED; }
	c->ctx.deferExpansion = on != 0;
	return SPR_OK;
}

SPR_API int sprContextSetAssemblyKernel(SprContext *c, uint32_t kernel)
{
	REQUIRE(c && kernel <= 2);
	c->ctx.assemblyKernel = kernel;
	return SPR_OK;
}

SPR_API int sprContextSetAssemblyShare(SprContext *c, uint32_t ctasPerSm)
{
	REQUIRE(c && ctasPerSm <= 8);
	c->ctx.assemblyCtasPerSm = ctasPerSm ? ctasPerSm : 8;
	return SPR_OK;
}

SPR_API int sprExpandRuns(SprContext *c, const SprGpuParticle *deviceRecords, uint64_t numRecords, SprGpuParticle *deviceVerticesOut)
{
	REQUIRE(c && deviceRecords && deviceVerticesOut);
	return guarded([&] {
		SPR_CUDA(cudaSetDevice(c->ctx.device));
		launch_expand_runs(c->ctx.stream, (const float4*)deviceRecords, numRecords, (float4*)deviceVerticesOut, c->ctx.numSms);
		c->ctx.launchCount++;
		SPR_CUDA(cudaGetLastError());
	});
}

}
