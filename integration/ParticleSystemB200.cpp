// ParticleSystemB200.cpp -- the literal drop-in.
//
// A second implementation of the reference's abstract class cl::ParticleSystem
// (src/client/ParticleSystem.h:9-22) whose methods forward to the C ABI of
// include/spear_particles.h. This file is compiled AGAINST THE REFERENCE'S HEADERS
// (it replaces src/client/ParticleSystem.cpp in the engine's build) and
// therefore only builds where the reference tree is available: oracle/build_ref.py builds it,
// together with the test harness, into oracle/_ref/libspear_adapter.so, and
// tests/test_gpu_adapter.py drives it through the reference's own types
// (sf::Box<sv::ParticleSystemComponent>, cl::Transform, cl::VisibleAreas, cl::FrameArgs,
// cl::RenderArgs) and checks it against the golden vectors of the reference implementation.

#include "client/ParticleSystem.h"

#include "client/AreaSystem.h"
#include "client/ParticleTexture.h"
#include "client/VisFogSystem.h"
#include "sf/Frustum.h"
#include "sp/Renderer.h"
#include "game/shader/GameShaders.h"
#include "game/shader/Particle.h"

#include <unordered_map>
#include <vector>
#include <memory>
#include <string.h>

#include "spear_particles.h"

namespace cl {

namespace {

// Flat POD mirror of one sv::ParticleSystemComponent (ServerState.h:187-239); the arrays behind the
// pointers live in the same object, so it can be handed to the C ABI for as long as the Box lives.
struct ComponentPod
{
	sf::Box<sv::ParticleSystemComponent> keepAlive;
	SprParticleSystemComponent pod;
	std::vector<SprVec2> splines[4];
	std::vector<SprGradientPoint> gradient;
	std::vector<SprGravityPoint> gravityPoints;
};

SprVec3 v3(const sf::Vec3 &v) { SprVec3 r = { v.x, v.y, v.z }; return r; }

void copyRandomVec3(SprRandomVec3 &d, const sv::RandomVec3 &s)
{
	d.offset = v3(s.offset); d.boxExtent = v3(s.boxExtent); d.rotation = v3(s.rotation);
	d.sphere.minTheta = s.sphere.minTheta; d.sphere.maxTheta = s.sphere.maxTheta;
	d.sphere.minPhi = s.sphere.minPhi; d.sphere.maxPhi = s.sphere.maxPhi;
	d.sphere.minRadius = s.sphere.minRadius; d.sphere.maxRadius = s.sphere.maxRadius;
	d.sphere.scale = v3(s.sphere.scale);
}

void copySpline(std::vector<SprVec2> &store, SprBSpline2 &d, const sv::BSpline2 &s)
{
	store.clear();
	for (const sf::Vec2 &p : s.points) { SprVec2 q = { p.x, p.y }; store.push_back(q); }
	d.points = store.data();
	d.numPoints = (uint32_t)store.size();
}

SprTransform convertTransform(const Transform &t)
{
	SprTransform r;
	r.position = v3(t.position);
	r.rotation.x = t.rotation.x; r.rotation.y = t.rotation.y; r.rotation.z = t.rotation.z; r.rotation.w = t.rotation.w;
	r.scale = t.scale;
	return r;
}

}

struct ParticleSystemB200 final : ParticleSystem
{
	SprContext *ctx = nullptr;
	SprSystem *sys = nullptr;
	std::unordered_map<const sv::ParticleSystemComponent*, std::unique_ptr<ComponentPod>> pods;
	std::vector<uint32_t> areaIds;
	std::vector<float> updateRadius;        // per effect: the descriptor's updateRadius (EffectType::updateRadius, ParticleSystem.cpp:303)
	std::vector<SprDrawRecord> lastRecords; // last renderMain, for tools that want to know which effect a draw was
	uint32_t lastNumSorted = 0;

	// The renderer side of the reference object (ParticleSystem.cpp:24-25, :253-263)
	static constexpr uint32_t MaxParticlesPerFrame = 4 * 1024;
	static constexpr uint32_t MaxParticleCacheFrames = 16;
	static constexpr uint32_t SplineSampleRate = 64, SplineAtlasWidth = 8, SplineAtlasHeight = 256, SplineCountPerType = 2;
	sp::Texture splineTexture;
	sp::Buffer vertexBuffers[MaxParticleCacheFrames];
	sp::Pipeline particlePipe;
	struct TypeSide { const void *owner = nullptr; sp::Ref<ParticleTexture> texture; }; // what initEffectType keeps outside the C ABI
	std::vector<TypeSide> typeSide;
	std::vector<uint32_t> uploadByteOffset; // per effect: Effect::uploadByteOffset (:196)
	std::vector<SprGpuParticle> staging;    // host copy of one effect's stream on its way into sg_append_buffer

	ParticleSystemB200(const SystemsDesc &desc)
	{
		SprContextDesc cd;
		memset(&cd, 0, sizeof(cd));
		cd.maxParticlesPerFrame = MaxParticlesPerFrame;
		cd.maxParticleCacheFrames = MaxParticleCacheFrames;
		sprContextCreate(&cd, &ctx);
		SprSystemsDesc sd;
		memset(&sd, 0, sizeof(sd));
		for (int i = 0; i < 4; i++) sd.seed[i] = desc.seed[i];
		sprSystemCreate(ctx, &sd, &sys);

		// the 16-deep ring of dynamic vertex buffers, the pipeline and the spline atlas, as the reference sets them up (:268-295)
		for (uint32_t i = 0; i < MaxParticleCacheFrames; i++) {
			sf::SmallStringBuf<128> name;
			name.format("particle vertexBuffer %u", i);
			vertexBuffers[i].initDynamicVertex(name.data, sizeof(SprGpuParticle) * 4 * MaxParticlesPerFrame);
		}
		{
			uint32_t flags = sp::PipeIndex16|sp::PipeDepthTest|sp::PipeBlendPremultiply;
			sg_pipeline_desc &d = particlePipe.init(gameShaders.particle, flags);
			d.blend.src_factor_alpha = SG_BLENDFACTOR_ZERO;
			d.blend.dst_factor_alpha = SG_BLENDFACTOR_ONE;
			d.layout.attrs[0].format = SG_VERTEXFORMAT_FLOAT4;
			d.layout.attrs[1].format = SG_VERTEXFORMAT_FLOAT4;
		}
		{
			sg_image_desc d = { };
			d.usage = SG_USAGE_IMMUTABLE;
			d.bqq_copy_target = true;
			d.pixel_format = SG_PIXELFORMAT_RGBA8;
			d.mag_filter = SG_FILTER_LINEAR;
			d.min_filter = SG_FILTER_LINEAR;
			d.width = SplineSampleRate * SplineAtlasWidth;
			d.height = SplineAtlasHeight * SplineCountPerType;
			d.label = "ParticleSystem splineTexture";
			splineTexture.init(d);
		}
	}

	~ParticleSystemB200()
	{
		if (sys) sprSystemDestroy(sys);
		if (ctx) sprContextDestroy(ctx);
	}

	const SprParticleSystemComponent *pod(const sf::Box<sv::ParticleSystemComponent> &c)
	{
		std::unique_ptr<ComponentPod> &slot = pods[c.ptr];
		if (!slot) {
			slot.reset(new ComponentPod());
			slot->keepAlive = c;
			const sv::ParticleSystemComponent &s = *c;
			SprParticleSystemComponent &d = slot->pod;
			sprComponentDefaults(&d);
			d.texture = (uint64_t)(uintptr_t)s.texture.data;
			d.frameCount[0] = s.frameCount.x; d.frameCount[1] = s.frameCount.y;
			d.timeStep = s.timeStep; d.prewarmTime = s.prewarmTime;
			d.updateRadius = s.updateRadius; d.cullPadding = s.cullPadding;
			d.updateOutOfCamera = s.updateOutOfCamera; d.renderOrder = s.renderOrder;
			d.spawnTime = s.spawnTime; d.spawnTimeVariance = s.spawnTimeVariance;
			d.burstAmount = s.burstAmount; d.burstAmountVariance = s.burstAmountVariance;
			d.emitterOnTime = s.emitterOnTime; d.localSpace = s.localSpace; d.instantDelete = s.instantDelete;
			copyRandomVec3(d.emitPosition, s.emitPosition);
			copyRandomVec3(d.emitVelocity, s.emitVelocity);
			d.emitVelocityAttractorOffset = v3(s.emitVelocityAttractorOffset);
			d.emitVelocityAttractorStrength = s.emitVelocityAttractorStrength;
			for (const sv::GravityPoint &p : s.gravityPoints) {
				SprGravityPoint q; q.position = v3(p.position); q.radius = p.radius; q.strength = p.strength;
				slot->gravityPoints.push_back(q);
			}
			d.gravityPoints = slot->gravityPoints.data();
			d.numGravityPoints = (uint32_t)slot->gravityPoints.size();
			d.drag = s.drag; d.gravity = v3(s.gravity);
			d.size = s.size; d.sizeVariance = s.sizeVariance;
			d.lifeTime = s.lifeTime; d.lifeTimeVariance = s.lifeTimeVariance;
			copySpline(slot->splines[0], d.scaleSpline, s.scaleSpline);
			copySpline(slot->splines[1], d.alphaSpline, s.alphaSpline);
			copySpline(slot->splines[2], d.additiveSpline, s.additiveSpline);
			copySpline(slot->splines[3], d.erosionSpline, s.erosionSpline);
			d.gradient.defaultColor = v3(s.gradient.defaultColor);
			for (const sv::GradientPoint &p : s.gradient.points) {
				SprGradientPoint q; q.t = p.t; q.color = v3(p.color);
				slot->gradient.push_back(q);
			}
			d.gradient.points = slot->gradient.data();
			d.gradient.numPoints = (uint32_t)slot->gradient.size();
			d.frameRate = s.frameRate; d.relativeFrameRate = s.relativeFrameRate; d.randomStartFrame = s.randomStartFrame;
			d.rotation = s.rotation; d.rotationVariance = s.rotationVariance; d.spin = s.spin; d.spinVariance = s.spinVariance;
		}
		return &slot->pod;
	}

	// The part of initEffectType that stays on the renderer's side (ParticleSystem.cpp:302, :421-449): the type's
	// texture reference and its two LUT rows copied into the spline atlas at the type's cell. The library bakes the
	// rows (sprGetEffectTypeLut); a type id is (re)initialised whenever another descriptor takes it over.
	void initTypeSide(uint32_t typeId, const sf::Box<sv::ParticleSystemComponent> &c)
	{
		if (typeSide.size() <= typeId) typeSide.resize(typeId + 1);
		TypeSide &side = typeSide[typeId];
		if (side.owner == (const void*)c.ptr) return;
		side.owner = (const void*)c.ptr;
		side.texture.load(c->texture);

		uint8_t splineTex[SPR_SPLINE_LUT_BYTES];
		sprGetEffectTypeLut(sys, typeId, splineTex);
		uint32_t atlasX = typeId / SplineAtlasHeight;
		uint32_t atlasY = typeId % SplineAtlasHeight;
		uint32_t x = atlasX * SplineSampleRate;
		uint32_t y = atlasY * SplineCountPerType;

		sg_image_desc d = { };
		d.width = SplineSampleRate;
		d.height = SplineCountPerType;
		d.pixel_format = SG_PIXELFORMAT_RGBA8;
		d.content.subimage[0][0].ptr = splineTex;
		d.content.subimage[0][0].size = sizeof(splineTex);
		sg_image stagingImage = sg_make_image(&d);
		{
			sg_bqq_subimage_rect rect;
			rect.src_x = 0;
			rect.src_y = 0;
			rect.dst_x = (int)x;
			rect.dst_y = (int)y;
			rect.dst_z = 0;
			rect.width = (int)SplineSampleRate;
			rect.height = (int)SplineCountPerType;

			sg_bqq_subimage_copy_desc desc;
			desc.dst_image = splineTexture.image;
			desc.src_image = stagingImage;
			desc.rects = &rect;
			desc.num_rects = 1;
			desc.num_mips = 1;
			sg_bqq_copy_subimages(&desc);
		}
		sg_destroy_image(stagingImage);
	}

	// -- cl::ParticleSystem

	virtual uint32_t reserveEffectType(Systems &, const sf::Box<sv::ParticleSystemComponent> &c) override
	{
		uint32_t typeId = 0;
		sprReserveEffectType(sys, pod(c), &typeId);
		initTypeSide(typeId, c);
		return typeId;
	}

	virtual void releaseEffectType(Systems &, const sf::Box<sv::ParticleSystemComponent> &c) override
	{
		sprReleaseEffectType(sys, pod(c));
	}

	virtual void addEffect(Systems &systems, uint32_t entityId, uint8_t componentIndex, const sf::Box<sv::ParticleSystemComponent> &c, const Transform &transform) override
	{
		SprTransform t = convertTransform(transform);
		uint32_t effectId = 0;
		sprAddEffect(sys, entityId, componentIndex, pod(c), &t, &effectId);
		{
			// addEffect reserves the effect's type itself (:739); while the effect holds that reference a reserve /
			// release pair only reads the id back
			uint32_t typeId = 0;
			sprReserveEffectType(sys, pod(c), &typeId);
			sprReleaseEffectType(sys, pod(c));
			initTypeSide(typeId, c);
		}
		if (updateRadius.size() <= effectId) { updateRadius.resize(effectId + 1, 0.0f); uploadByteOffset.resize(effectId + 1, 0); }
		updateRadius[effectId] = c->updateRadius;
		uploadByteOffset[effectId] = 0;
		sf::Sphere sphere = { transform.position, c->updateRadius };
		uint32_t areaId = systems.area->addSphereArea(AreaGroup::ParticleEffect, effectId, sphere, Area::Activate | Area::Visibility);
		if (areaIds.size() <= effectId) areaIds.resize(effectId + 1);
		areaIds[effectId] = areaId;
		systems.entities.addComponent(entityId, this, effectId, 0, componentIndex, Entity::UpdateTransform|Entity::PrepareForRemove);
	}

	virtual void updateParticles(const VisibleAreas &activeAreas, const FrameArgs &args) override
	{
		sf::Slice<const uint32_t> ids = activeAreas.get(AreaGroup::ParticleEffect);
		SprFrameArgs fa;
		fa.frameIndex = args.frameIndex; fa.gameTime = args.gameTime; fa.dt = args.dt;
		fa.cameraPosition = v3(args.mainRenderArgs.cameraPosition);
		sprUpdateParticles(sys, ids.data, (uint32_t)ids.size, &fa);
	}

	virtual void renderMain(const VisFogSystem *visFogSystem, const VisibleAreas &visibleAreas, const RenderArgs &renderArgs) override
	{
		VisFogImage visFog = visFogSystem->getVisFogImage();
		SprRenderArgs ra;
		memset(&ra, 0, sizeof(ra));
		ra.cameraPosition = v3(renderArgs.cameraPosition);
		memcpy(ra.viewToClip.v, renderArgs.viewToClip.v, sizeof(float) * 16);
		memcpy(ra.worldToClip.v, renderArgs.worldToClip.v, sizeof(float) * 16);
		for (int i = 0; i < 4; i++) {
			renderArgs.frustum.side[i].storeu(ra.frustum.side[i]);
			renderArgs.frustum.caps[i].storeu(ra.frustum.caps[i]);
		}
		ra.visFogWorldMad.x = visFog.worldMad.x; ra.visFogWorldMad.y = visFog.worldMad.y;
		ra.visFogWorldMad.z = visFog.worldMad.z; ra.visFogWorldMad.w = visFog.worldMad.w;

		sf::Slice<const uint32_t> ids = visibleAreas.get(AreaGroup::ParticleEffect);
		SprRenderOutput out;
		if (sprRenderMain(sys, ids.data, (uint32_t)ids.size, &ra, &out) != SPR_OK) return;
		lastRecords.assign(out.records, out.records + out.numRecords);
		lastNumSorted = out.numSortedEffects;

		for (uint32_t i = 0; i < out.numRecords; i++) {
			const SprDrawRecord &d = out.records[i];

			// The library has made the reference's upload decision (cache age and the per-frame budget, :861-863): an
			// effect whose cached copy is stale comes back to the host -- at most MaxParticlesPerFrame particles a frame,
			// 512 KB -- and goes into this frame's ring buffer exactly as the reference's own array does (:865).
			if (d.uploadedThisFrame) {
				staging.resize((size_t)d.numParticles * 4);
				uint32_t numVertices = 0;
				sprReadVertexStream(sys, d.effectId, staging.data(), (uint32_t)staging.size(), &numVertices);
				uploadByteOffset[d.effectId] = (uint32_t)sg_append_buffer(vertexBuffers[d.uploadRingSlot].buffer, staging.data(), (int)(staging.size() * sizeof(SprGpuParticle)));
			}
			sp::Buffer &vertexBuffer = vertexBuffers[d.uploadRingSlot];

			particlePipe.bind();
			if (d.typeUboChanged) {
				SprVertexTypeUbo typeUbo;
				sprGetEffectTypeUbo(sys, d.typeId, &typeUbo);
				sg_apply_uniforms(SG_SHADERSTAGE_VS, SLOT_Particle_VertexType, &typeUbo, sizeof(typeUbo));
			}
			sg_apply_uniforms(SG_SHADERSTAGE_VS, SLOT_Particle_VertexInstance, &d.instanceUbo, sizeof(d.instanceUbo));

			sg_image image = ParticleTexture::defaultImage;
			const TypeSide &side = typeSide[d.typeId];
			if (side.texture.isLoaded() && side.texture->image.id) {
				image = side.texture->image;
			}

			sg_bindings binds = { };
			binds.vertex_buffers[0] = vertexBuffer.buffer;
			binds.vertex_buffer_offsets[0] = (int)uploadByteOffset[d.effectId];
			binds.index_buffer = sp::getSharedQuadIndexBuffer();
			binds.vs_images[SLOT_Particle_u_SplineTexture] = splineTexture.image;
			binds.vs_images[SLOT_Particle_u_VisFogTexture] = visFog.image;
			binds.fs_images[SLOT_Particle_u_Texture] = image;
			sg_apply_bindings(&binds);
			sg_draw(0, (int)d.numParticles * 6, 1);
		}
	}

	// -- cl::EntitySystem

	void updateTransform(Systems &systems, uint32_t, const EntityComponent &ec, const TransformUpdate &update) override
	{
		SprTransformUpdate u;
		u.previousTransform = convertTransform(update.previousTransform);
		u.transform = convertTransform(update.transform);
		memcpy(u.entityToWorld.v, update.entityToWorld.v, sizeof(float) * 12);
		sprUpdateTransform(sys, ec.userId, &u);
		// the activation / visibility sphere follows the emitter (ParticleSystem.cpp:782-783)
		sf::Sphere sphere = { update.transform.position, ec.userId < updateRadius.size() ? updateRadius[ec.userId] : 0.0f };
		if (ec.userId < areaIds.size()) systems.area->updateSphereArea(areaIds[ec.userId], sphere);
	}

	bool prepareForRemove(Systems &, uint32_t, const EntityComponent &ec, const FrameArgs &frameArgs) override
	{
		SprFrameArgs fa;
		memset(&fa, 0, sizeof(fa));
		fa.frameIndex = frameArgs.frameIndex; fa.gameTime = frameArgs.gameTime; fa.dt = frameArgs.dt;
		int canRemove = 0;
		sprPrepareForRemove(sys, ec.userId, &fa, &canRemove);
		return canRemove != 0;
	}

	void remove(Systems &systems, uint32_t, const EntityComponent &ec) override
	{
		if (ec.userId < areaIds.size()) systems.area->removeSphereArea(areaIds[ec.userId]);
		sprRemoveEffect(sys, ec.userId);
	}
};

sf::Box<ParticleSystem> ParticleSystem::create(const SystemsDesc &desc) { return sf::box<ParticleSystemB200>(desc); }

// accessors for the test harness (oracle/ref_harness.cpp built with -DSPR_ADAPTER)
SprSystem *sprAdapterSystem(ParticleSystem *ps) { return ((ParticleSystemB200*)ps)->sys; }
SprContext *sprAdapterContext(ParticleSystem *ps) { return ((ParticleSystemB200*)ps)->ctx; }
const SprDrawRecord *sprAdapterLastRecords(ParticleSystem *ps, uint32_t *numSorted)
{
	ParticleSystemB200 *p = (ParticleSystemB200*)ps;
	if (numSorted) *numSorted = p->lastNumSorted;
	return p->lastRecords.data();
}

}
