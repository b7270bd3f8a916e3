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
	sp::Pipeline particlePipe;
	std::vector<SprDrawRecord> lastRecords; // last renderMain, for tools that want to know which effect a draw was
	uint32_t lastNumSorted = 0;

	ParticleSystemB200(const SystemsDesc &desc)
	{
		SprContextDesc cd;
		memset(&cd, 0, sizeof(cd));   // defaults: 4096 particles per frame, 16 cache frames (ParticleSystem.cpp:24-25)
		sprContextCreate(&cd, &ctx);
		SprSystemsDesc sd;
		memset(&sd, 0, sizeof(sd));
		for (int i = 0; i < 4; i++) sd.seed[i] = desc.seed[i];
		sprSystemCreate(ctx, &sd, &sys);
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

	// -- cl::ParticleSystem

	virtual uint32_t reserveEffectType(Systems &, const sf::Box<sv::ParticleSystemComponent> &c) override
	{
		uint32_t typeId = 0;
		sprReserveEffectType(sys, pod(c), &typeId);
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
			particlePipe.bind();
			if (d.typeUboChanged) {
				SprVertexTypeUbo typeUbo;
				sprGetEffectTypeUbo(sys, d.typeId, &typeUbo);
				sg_apply_uniforms(SG_SHADERSTAGE_VS, SLOT_Particle_VertexType, &typeUbo, sizeof(typeUbo));
			}
			sg_apply_uniforms(SG_SHADERSTAGE_VS, SLOT_Particle_VertexInstance, &d.instanceUbo, sizeof(d.instanceUbo));
			// The vertices live in GPU memory at out.deviceVertices + d.vertexByteOffset. A real renderer binds
			// that range through CUDA-graphics interop; the sokol bindings below carry the offset only.
			sg_bindings binds = { };
			binds.vertex_buffer_offsets[0] = (int)d.vertexByteOffset;
			binds.index_buffer = sp::getSharedQuadIndexBuffer();
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
		// updateRadius is a property of the descriptor; the area system keeps the sphere (ParticleSystem.cpp:782-783)
		(void)systems;
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
