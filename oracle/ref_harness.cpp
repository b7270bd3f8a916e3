// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (see oracle_api.h).
//
// Headless harness around the UNMODIFIED reference hot path. This TU includes
// /root/reference/src/client/ParticleSystem.cpp (found through the include path
// set by oracle/build_ref.py) so that it can reach ParticleSystemImp's internals,
// and supplies link stubs for the renderer / asset / entity services that file
// touches. No reference source is copied into this repo:
// the reference files are compiled where they lie and only the resulting
// library lands in oracle/_ref/.

#ifdef SPR_ADAPTER
// Adapter variant (kind "adapter"): the same harness drives integration/ParticleSystemB200.cpp, the
// cl::ParticleSystem implementation that forwards to libspear_particles.so, through the reference's
// own types; read-backs go through the C ABI instead of the reference's internals.
#include "client/ParticleSystem.h"
#include "client/VisFogSystem.h"
#include "client/ParticleTexture.h"
#include "sf/Frustum.h"
#include "sf/Float4.h"
#include "sp/Asset.h"
#include "sp/Renderer.h"
#include "game/shader/GameShaders.h"
#include "game/shader/Particle.h"
#include "spear_particles.h"
namespace cl { SprSystem *sprAdapterSystem(ParticleSystem *ps); SprContext *sprAdapterContext(ParticleSystem *ps);
	const SprDrawRecord *sprAdapterLastRecords(ParticleSystem *ps, uint32_t *numSorted); }
#else
#include "client/ParticleSystem.cpp"
#endif

#include "client/AreaSystem.h"
#include "sf/Geometry.h"

#include <chrono>
#include <map>
#include <unordered_map>
#include <vector>
#include <string.h>

#include "oracle_api.h"

// ---------------------------------------------------------------------------
// Link stubs (definitions only, no behaviour beyond what the harness captures)
// ---------------------------------------------------------------------------

GameShaders gameShaders;
void GameShaders::load() { }

namespace sp {

sg_buffer g_hackSharedQuadIndexBuffer = { 7777 };

Buffer::Buffer() { }
Buffer::~Buffer() { }
void Buffer::reset() { }
// dynamic vertex buffers get the ids 1000, 1001, ... in creation order: a system's ring (ParticleSystem.cpp:268-272)
// is [first id of the system, +16)
static uint32_t g_nextVertexBufferId = 1000;
void Buffer::initDynamicVertex(const char *, size_t) { buffer.id = g_nextVertexBufferId++; }

Pipeline::Pipeline() { }
Pipeline::~Pipeline() { }
sg_pipeline_desc &Pipeline::init(sg_shader, uint32_t) { return desc; }
bool Pipeline::bind() { return true; }
void Pipeline::reset() { }

Texture::Texture() { }
Texture::~Texture() { }
void Texture::init(const sg_image_desc &) { image.id = 1; }
void Texture::reset() { }

AssetProps::~AssetProps() { }
uint32_t NoAssetProps::hash() const { return 0; }
bool NoAssetProps::equal(const AssetProps &) const { return true; }
void NoAssetProps::copyTo(AssetProps *) const { }

Asset::Asset() { }
Asset::~Asset() { }
Asset *Asset::impCreate(AssetType *, const sf::Symbol &, const AssetProps &) { return nullptr; }
bool Asset::isLoaded() const { return false; }
void Asset::release() { }
void Asset::retain() { }

}

namespace cl {

sp::AssetType ParticleTexture::SelfType;
sg_image ParticleTexture::defaultImage;

System::~System() { }
bool EntitySystem::prepareForRemove(Systems &, uint32_t, const EntityComponent &, const FrameArgs &) { return true; }
void EntitySystem::editorHighlight(Systems &, const EntityComponent &, EditorHighlight) { }

static uint32_t g_lastAddedEffectId;
void Entities::addComponent(uint32_t, EntitySystem *, uint32_t userId, uint8_t, uint8_t, uint32_t)
{
	g_lastAddedEffectId = userId;
}

}

// -- sokol_gfx capture stubs

static uint8_t g_lastLut[SPR_SPLINE_LUT_BYTES];
static bool g_lutCaptured;

struct CapturedDraw
{
	SprVertexInstanceUbo instanceUbo;
	uint32_t typeUboChanged;
	uint32_t numElements;
	uint32_t vertexBuffer;
	uint32_t vertexOffset;
	uint32_t indexBuffer, vsImage0, vsImage1, fsImage0;
};
static std::vector<CapturedDraw> g_draws;
static CapturedDraw g_curDraw;
// what sg_append_buffer was given during the last renderMain (ParticleSystem.cpp:865)
struct CapturedAppend { uint32_t buffer, offset, numBytes; uint64_t hash; };
static std::vector<CapturedAppend> g_appends;
static std::unordered_map<uint32_t, uint32_t> g_appendOffsets;
// the spline atlas as the sg_make_image / sg_bqq_copy_subimages calls leave it (:253-258, :421-449): 512 x 512 RGBA8
static const uint32_t kAtlasRes = 512;
static std::vector<uint8_t> *g_curAtlas;     // atlas of the system whose call is in progress
static std::vector<uint8_t> g_stagingContent; // content of the last sg_make_image
static uint32_t g_stagingW, g_stagingH;

static uint64_t fnv1a64(const void *data, size_t n)
{
	uint64_t h = 0xcbf29ce484222325ull;
	const uint8_t *p = (const uint8_t*)data;
	for (size_t i = 0; i < n; i++) h = (h ^ p[i]) * 0x100000001b3ull;
	return h;
}

extern "C" {

sg_image sg_make_image(const sg_image_desc *desc)
{
	if (desc->content.subimage[0][0].ptr && desc->content.subimage[0][0].size == SPR_SPLINE_LUT_BYTES) {
		memcpy(g_lastLut, desc->content.subimage[0][0].ptr, SPR_SPLINE_LUT_BYTES);
		g_lutCaptured = true;
	}
	g_stagingContent.clear();
	if (desc->content.subimage[0][0].ptr) {
		const uint8_t *p = (const uint8_t*)desc->content.subimage[0][0].ptr;
		g_stagingContent.assign(p, p + desc->content.subimage[0][0].size);
	}
	g_stagingW = (uint32_t)desc->width; g_stagingH = (uint32_t)desc->height;
	sg_image img; img.id = 2; return img;
}
void sg_destroy_image(sg_image) { }
void sg_bqq_copy_subimages(const sg_bqq_subimage_copy_desc *d)
{
	if (!g_curAtlas || d->dst_image.id != 1 || d->src_image.id != 2) return;
	if (g_curAtlas->empty()) g_curAtlas->assign((size_t)kAtlasRes * kAtlasRes * 4, 0);
	for (int r = 0; r < d->num_rects; r++) {
		const sg_bqq_subimage_rect &rc = d->rects[r];
		for (int y = 0; y < rc.height; y++) {
			for (int x = 0; x < rc.width; x++) {
				size_t src = ((size_t)(rc.src_y + y) * g_stagingW + (size_t)(rc.src_x + x)) * 4;
				size_t dst = ((size_t)(rc.dst_y + y) * kAtlasRes + (size_t)(rc.dst_x + x)) * 4;
				if (src + 4 <= g_stagingContent.size() && dst + 4 <= g_curAtlas->size()) memcpy(g_curAtlas->data() + dst, g_stagingContent.data() + src, 4);
			}
		}
	}
}
int sg_append_buffer(sg_buffer buf, const void *data, int numBytes)
{
	uint32_t &off = g_appendOffsets[buf.id];
	int r = (int)off;
	off += (uint32_t)numBytes;
	g_appends.push_back({ buf.id, (uint32_t)r, (uint32_t)numBytes, fnv1a64(data, (size_t)numBytes) });
	return r;
}
void sg_apply_bindings(const sg_bindings *b)
{
	g_curDraw.vertexBuffer = b->vertex_buffers[0].id;
	g_curDraw.vertexOffset = (uint32_t)b->vertex_buffer_offsets[0];
	g_curDraw.indexBuffer = b->index_buffer.id;
	g_curDraw.vsImage0 = b->vs_images[SLOT_Particle_u_SplineTexture].id;
	g_curDraw.vsImage1 = b->vs_images[SLOT_Particle_u_VisFogTexture].id;
	g_curDraw.fsImage0 = b->fs_images[SLOT_Particle_u_Texture].id;
}
void sg_apply_uniforms(sg_shader_stage, int slot, const void *data, int numBytes)
{
	if (slot == SLOT_Particle_VertexInstance && numBytes == (int)sizeof(SprVertexInstanceUbo)) {
		memcpy(&g_curDraw.instanceUbo, data, sizeof(SprVertexInstanceUbo));
	} else if (slot == SLOT_Particle_VertexType) {
		g_curDraw.typeUboChanged = 1;
	}
}
void sg_draw(int, int numElements, int)
{
	g_curDraw.numElements = (uint32_t)numElements;
	g_draws.push_back(g_curDraw);
	memset(&g_curDraw, 0, sizeof(g_curDraw));
}

}

static_assert(sizeof(SprVertexInstanceUbo) == sizeof(Particle_VertexInstance_t), "instance ubo size");
static_assert(sizeof(SprVertexTypeUbo) == sizeof(Particle_VertexType_t), "type ubo size");
static_assert(sizeof(SprGpuParticle) == 32, "gpu particle size");

// ---------------------------------------------------------------------------
// Service stubs with behaviour the path needs
// ---------------------------------------------------------------------------

namespace {

struct StubAreaSystem final : cl::AreaSystem
{
	uint32_t nextId = 0;
	uint32_t addBoxArea(cl::AreaGroup, uint32_t, const sf::Bounds3 &, uint32_t) override { return nextId++; }
	uint32_t addBoxArea(cl::AreaGroup, uint32_t, const sf::Bounds3 &, const sf::Mat34 &, uint32_t) override { return nextId++; }
	void updateBoxArea(uint32_t, const sf::Bounds3 &) override { }
	void updateBoxArea(uint32_t, const sf::Bounds3 &, const sf::Mat34 &) override { }
	void removeBoxArea(uint32_t) override { }
	// the sphere areas the particle system registers and keeps current (ParticleSystem.cpp:762-763, :782-783, :801)
	struct SphereArea { uint32_t userId; float x, y, z, r; };
	std::map<uint32_t, SphereArea> spheres;
	uint32_t addSphereArea(cl::AreaGroup, uint32_t userId, const sf::Sphere &s, uint32_t) override
	{ spheres[nextId] = { userId, s.origin.x, s.origin.y, s.origin.z, s.radius }; return nextId++; }
	uint32_t addSphereArea(cl::AreaGroup, uint32_t, const sf::Sphere &, const sf::Mat34 &, uint32_t) override { return nextId++; }
	void updateSphereArea(uint32_t areaId, const sf::Sphere &s) override
	{ auto it = spheres.find(areaId); if (it != spheres.end()) { it->second.x = s.origin.x; it->second.y = s.origin.y; it->second.z = s.origin.z; it->second.r = s.radius; } }
	void updateSphereArea(uint32_t, const sf::Sphere &, const sf::Mat34 &) override { }
	void removeSphereArea(uint32_t areaId) override { spheres.erase(areaId); }
	void optimize() override { }
	void queryFrustum(sf::Array<cl::Area> &, uint32_t, const sf::Frustum &) const override { }
	void queryFrustumBounds(sf::Array<cl::AreaBounds> &, uint32_t, const sf::Frustum &) const override { }
	void castRay(sf::Array<cl::Area> &, uint32_t, const sf::FastRay &, float, float) const override { }
};

struct StubVisFog final : cl::VisFogSystem
{
	sf::Vec4 worldMad;
	void updateVisibility(const sv::VisibleUpdateEvent &, bool) override { }
	void disableForFrame() override { }
	void updateTexture(float) override { }
	cl::VisFogImage getVisFogImage() const override { cl::VisFogImage img; img.worldMad = worldMad; img.image.id = 3; return img; }
};

struct RefSystem
{
	cl::Systems systems;
	sf::Box<cl::ParticleSystem> ps;
#ifdef SPR_ADAPTER
	SprSystem *spr = nullptr;
	std::vector<uint32_t> lastSorted;
#else
	cl::ParticleSystemImp *imp = nullptr;
#endif
	std::unordered_map<const SprParticleSystemComponent*, sf::Box<sv::ParticleSystemComponent>> comps;
	std::unordered_map<uint32_t, std::vector<uint8_t>> luts;
	std::vector<uint8_t> atlas;          // spline atlas texels as the sg_* calls of this system left them
	uint32_t firstVertexBuffer = 0;      // id of ring buffer 0 of this system
	std::vector<uint64_t> simSteps;
	cl::VisibleAreas areas;
	StubVisFog visFog;
};

sf::Vec3 v3(const SprVec3 &v) { return sf::Vec3(v.x, v.y, v.z); }

void convertRandomVec3(sv::RandomVec3 &d, const SprRandomVec3 &s)
{
	d.offset = v3(s.offset);
	d.boxExtent = v3(s.boxExtent);
	d.sphere.minTheta = s.sphere.minTheta; d.sphere.maxTheta = s.sphere.maxTheta;
	d.sphere.minPhi = s.sphere.minPhi; d.sphere.maxPhi = s.sphere.maxPhi;
	d.sphere.minRadius = s.sphere.minRadius; d.sphere.maxRadius = s.sphere.maxRadius;
	d.sphere.scale = v3(s.sphere.scale);
	d.rotation = v3(s.rotation);
}

void convertSpline(sv::BSpline2 &d, const SprBSpline2 &s)
{
	d.points.clear();
	for (uint32_t i = 0; i < s.numPoints; i++) d.points.push(sf::Vec2(s.points[i].x, s.points[i].y));
}

sf::Box<sv::ParticleSystemComponent> convertComponent(const SprParticleSystemComponent &s)
{
	sf::Box<sv::ParticleSystemComponent> box = sf::box<sv::ParticleSystemComponent>();
	sv::ParticleSystemComponent &d = *box;
	d.frameCount = sf::Vec2i(s.frameCount[0], s.frameCount[1]);
	d.timeStep = s.timeStep; d.prewarmTime = s.prewarmTime;
	d.updateRadius = s.updateRadius; d.cullPadding = s.cullPadding;
	d.updateOutOfCamera = s.updateOutOfCamera != 0; d.renderOrder = s.renderOrder;
	d.spawnTime = s.spawnTime; d.spawnTimeVariance = s.spawnTimeVariance;
	d.burstAmount = s.burstAmount; d.burstAmountVariance = s.burstAmountVariance;
	d.emitterOnTime = s.emitterOnTime;
	d.localSpace = s.localSpace != 0; d.instantDelete = s.instantDelete != 0;
	convertRandomVec3(d.emitPosition, s.emitPosition);
	convertRandomVec3(d.emitVelocity, s.emitVelocity);
	d.emitVelocityAttractorOffset = v3(s.emitVelocityAttractorOffset);
	d.emitVelocityAttractorStrength = s.emitVelocityAttractorStrength;
	for (uint32_t i = 0; i < s.numGravityPoints; i++) {
		sv::GravityPoint &p = d.gravityPoints.push();
		p.position = v3(s.gravityPoints[i].position);
		p.radius = s.gravityPoints[i].radius;
		p.strength = s.gravityPoints[i].strength;
	}
	d.drag = s.drag; d.gravity = v3(s.gravity);
	d.size = s.size; d.sizeVariance = s.sizeVariance;
	d.lifeTime = s.lifeTime; d.lifeTimeVariance = s.lifeTimeVariance;
	convertSpline(d.scaleSpline, s.scaleSpline);
	convertSpline(d.alphaSpline, s.alphaSpline);
	convertSpline(d.additiveSpline, s.additiveSpline);
	convertSpline(d.erosionSpline, s.erosionSpline);
	d.gradient.defaultColor = v3(s.gradient.defaultColor);
	for (uint32_t i = 0; i < s.gradient.numPoints; i++) {
		sv::GradientPoint &p = d.gradient.points.push();
		p.t = s.gradient.points[i].t;
		p.color = v3(s.gradient.points[i].color);
	}
	d.frameRate = s.frameRate;
	d.relativeFrameRate = s.relativeFrameRate != 0; d.randomStartFrame = s.randomStartFrame != 0;
	d.rotation = s.rotation; d.rotationVariance = s.rotationVariance;
	d.spin = s.spin; d.spinVariance = s.spinVariance;
	return box;
}

const sf::Box<sv::ParticleSystemComponent> &getComponent(RefSystem *rs, const SprParticleSystemComponent *c)
{
	auto it = rs->comps.find(c);
	if (it == rs->comps.end()) {
		it = rs->comps.emplace(c, convertComponent(*c)).first;
	}
	return it->second;
}

cl::Transform convertTransform(const SprTransform &t)
{
	cl::Transform r;
	r.position = v3(t.position);
	r.rotation = sf::Quat(t.rotation.x, t.rotation.y, t.rotation.z, t.rotation.w);
	r.scale = t.scale;
	return r;
}

void setActive(RefSystem *rs, const uint32_t *ids, uint32_t n)
{
	sf::Array<uint32_t> &g = rs->areas.groups[(uint32_t)cl::AreaGroup::ParticleEffect];
	g.clear();
	for (uint32_t i = 0; i < n; i++) g.push(ids[i]);
}

cl::FrameArgs convertFrameArgs(const SprFrameArgs &a)
{
	cl::FrameArgs f;
	f.frameIndex = a.frameIndex;
	f.gameTime = a.gameTime;
	f.dt = a.dt;
	f.mainRenderArgs.cameraPosition = v3(a.cameraPosition);
	return f;
}

#ifdef SPR_ADAPTER
void countSteps(RefSystem *, const uint32_t *, uint32_t, float) { } // the library counts sim steps itself
#else
void countSteps(RefSystem *rs, const uint32_t *ids, uint32_t n, float dt)
{
	float clampedDt = sf::min(dt, 0.1f);
	for (uint32_t i = 0; i < n; i++) {
		uint32_t id = ids[i];
		if (rs->simSteps.size() <= id) rs->simSteps.resize(id + 1, 0);
		const auto &e = rs->imp->effects[id];
		volatile float td = e.timeDelta;
		td = td + clampedDt;
		while (td >= e.timeStep) { rs->simSteps[id]++; td = td - e.timeStep; }
	}
}
#endif

}

// ---------------------------------------------------------------------------
// orc_* API
// ---------------------------------------------------------------------------

extern "C" {

#ifdef SPR_ADAPTER
const char *orc_kind(void) { return "adapter"; }
#else
const char *orc_kind(void) { return "reference"; }
#endif

void *orc_create(const SprSystemsDesc *desc)
{
	RefSystem *rs = new RefSystem();
	cl::SystemsDesc d;
	d.persist.camera = sf::Vec2(desc->persistCamera[0], desc->persistCamera[1]);
	d.persist.zoom = desc->persistZoom;
	for (int i = 0; i < 4; i++) d.seed[i] = desc->seed[i];
	rs->systems.area = sf::box<StubAreaSystem>();
	rs->firstVertexBuffer = sp::g_nextVertexBufferId;
	g_curAtlas = &rs->atlas;
	rs->ps = cl::ParticleSystem::create(d);
#ifdef SPR_ADAPTER
	rs->spr = cl::sprAdapterSystem(rs->ps.ptr);
	if (!rs->spr) { delete rs; return nullptr; }
#else
	rs->imp = (cl::ParticleSystemImp*)rs->ps.ptr;
#endif
	rs->systems.particle = rs->ps;
	return rs;
}

void orc_destroy(void *sys) { if (g_curAtlas == &((RefSystem*)sys)->atlas) g_curAtlas = nullptr; delete (RefSystem*)sys; }

uint32_t orc_reserve_type(void *sys, const SprParticleSystemComponent *c)
{
	RefSystem *rs = (RefSystem*)sys;
	g_lutCaptured = false;
	g_curAtlas = &rs->atlas;
	uint32_t typeId = rs->ps->reserveEffectType(rs->systems, getComponent(rs, c));
	if (g_lutCaptured) rs->luts[typeId] = std::vector<uint8_t>(g_lastLut, g_lastLut + SPR_SPLINE_LUT_BYTES);
	return typeId;
}

void orc_release_type(void *sys, const SprParticleSystemComponent *c)
{
	RefSystem *rs = (RefSystem*)sys;
	rs->ps->releaseEffectType(rs->systems, getComponent(rs, c));
}

uint32_t orc_add_effect(void *sys, uint32_t entityId, uint8_t componentIndex,
	const SprParticleSystemComponent *c, const SprTransform *transform)
{
	RefSystem *rs = (RefSystem*)sys;
	g_lutCaptured = false;
	g_curAtlas = &rs->atlas;
	rs->ps->addEffect(rs->systems, entityId, componentIndex, getComponent(rs, c), convertTransform(*transform));
	uint32_t effectId = cl::g_lastAddedEffectId;
#ifndef SPR_ADAPTER
	if (g_lutCaptured) rs->luts[rs->imp->effects[effectId].typeId] = std::vector<uint8_t>(g_lastLut, g_lastLut + SPR_SPLINE_LUT_BYTES);
#endif
	if (rs->simSteps.size() <= effectId) rs->simSteps.resize(effectId + 1, 0);
	rs->simSteps[effectId] = 0;
	return effectId;
}

void orc_update_transform(void *sys, uint32_t effectId, const SprTransformUpdate *update)
{
	RefSystem *rs = (RefSystem*)sys;
	cl::TransformUpdate u;
	u.previousTransform = convertTransform(update->previousTransform);
	u.transform = convertTransform(update->transform);
	memcpy(u.entityToWorld.v, update->entityToWorld.v, sizeof(float) * 12);
	cl::EntityComponent ec = { };
	ec.system = rs->ps.ptr;
	ec.userId = effectId;
	((cl::EntitySystem*)rs->ps.ptr)->updateTransform(rs->systems, 0, ec, u);
}

int orc_prepare_for_remove(void *sys, uint32_t effectId, const SprFrameArgs *args)
{
	RefSystem *rs = (RefSystem*)sys;
	cl::EntityComponent ec = { };
	ec.system = rs->ps.ptr;
	ec.userId = effectId;
	return ((cl::EntitySystem*)rs->ps.ptr)->prepareForRemove(rs->systems, 0, ec, convertFrameArgs(*args)) ? 1 : 0;
}

void orc_remove(void *sys, uint32_t effectId)
{
	RefSystem *rs = (RefSystem*)sys;
	cl::EntityComponent ec = { };
	ec.system = rs->ps.ptr;
	ec.userId = effectId;
	((cl::EntitySystem*)rs->ps.ptr)->remove(rs->systems, 0, ec);
}

void orc_update(void *sys, const uint32_t *activeIds, uint32_t numActive, const SprFrameArgs *args)
{
	RefSystem *rs = (RefSystem*)sys;
	setActive(rs, activeIds, numActive);
	countSteps(rs, activeIds, numActive, args->dt);
	rs->ps->updateParticles(rs->areas, convertFrameArgs(*args));
}

uint32_t orc_render(void *sys, const uint32_t *visibleIds, uint32_t numVisible,
	const SprRenderArgs *args, SprDrawRecord *out, uint32_t capacity, uint32_t *numSorted)
{
	RefSystem *rs = (RefSystem*)sys;
	setActive(rs, visibleIds, numVisible);
	cl::RenderArgs ra;
	ra.cameraPosition = v3(args->cameraPosition);
	memcpy(ra.viewToClip.v, args->viewToClip.v, sizeof(float) * 16);
	memcpy(ra.worldToClip.v, args->worldToClip.v, sizeof(float) * 16);
	for (int i = 0; i < 4; i++) {
		ra.frustum.side[i] = sf::Float4::loadu(args->frustum.side[i]);
		ra.frustum.caps[i] = sf::Float4::loadu(args->frustum.caps[i]);
	}
	rs->visFog.worldMad = sf::Vec4(args->visFogWorldMad.x, args->visFogWorldMad.y, args->visFogWorldMad.z, args->visFogWorldMad.w);
	g_draws.clear();
	g_appends.clear();
	memset(&g_curDraw, 0, sizeof(g_curDraw));
	rs->ps->renderMain(&rs->visFog, rs->areas, ra);
#ifdef SPR_ADAPTER
	uint32_t nSorted = 0;
	const SprDrawRecord *last = cl::sprAdapterLastRecords(rs->ps.ptr, &nSorted);
	if (numSorted) *numSorted = nSorted;
	uint32_t n = 0;
	for (size_t i = 0; i < g_draws.size() && n < capacity; i++, n++) {
		const CapturedDraw &d = g_draws[i];
		SprDrawRecord &r = out[n];
		memset(&r, 0, sizeof(r));
		r.effectId = last[i].effectId;           // which effect a draw belongs to is not visible in the sg_* calls
		r.typeId = last[i].typeId;
		r.numParticles = d.numElements / 6;      // everything else is what the adapter really submitted
		r.typeUboChanged = d.typeUboChanged;
		r.vertexByteOffset = d.vertexOffset;
		r.instanceUbo = d.instanceUbo;
	}
	return n;
}

void orc_effect_info(void *sys, uint32_t effectId, SprEffectInfo *out) { sprGetEffectInfo(((RefSystem*)sys)->spr, effectId, out); }

uint32_t orc_read_free(void *sys, uint32_t effectId, uint32_t *out, uint32_t capacity)
{
	uint32_t n = 0; sprReadFreeIndices(((RefSystem*)sys)->spr, effectId, out, capacity, &n); return n;
}

uint32_t orc_read_stream(void *sys, uint32_t effectId, SprGpuParticle *out, uint32_t capacityVertices)
{
	uint32_t n = 0; sprReadVertexStream(((RefSystem*)sys)->spr, effectId, out, capacityVertices, &n); return n;
}

uint32_t orc_read_slots(void *sys, uint32_t effectId, SprGpuParticle *out, uint32_t capacitySlots)
{
	uint32_t n = 0; sprReadSlots(((RefSystem*)sys)->spr, effectId, out, capacitySlots, &n); return n;
}

void orc_rng_state(void *sys, uint64_t outStateInc[2]) { sprGetRngState(((RefSystem*)sys)->spr, outStateInc); }
void orc_type_lut(void *sys, uint32_t typeId, uint8_t out[SPR_SPLINE_LUT_BYTES]) { sprGetEffectTypeLut(((RefSystem*)sys)->spr, typeId, out); }
void orc_type_ubo(void *sys, uint32_t typeId, SprVertexTypeUbo *out) { sprGetEffectTypeUbo(((RefSystem*)sys)->spr, typeId, out); }

#else
	if (numSorted) *numSorted = rs->imp->sortedEffects.size;
	uint32_t n = 0;
	// Draw k corresponds to sortedEffects[k] (the loop at ParticleSystem.cpp:851 only
	// skips entries that the build loop at :835-846 already skipped).
	for (size_t i = 0; i < g_draws.size() && n < capacity; i++, n++) {
		const CapturedDraw &d = g_draws[i];
		SprDrawRecord &r = out[n];
		memset(&r, 0, sizeof(r));
		r.effectId = rs->imp->sortedEffects[(uint32_t)i].effectId;
		r.typeId = rs->imp->effects[r.effectId].typeId;
		r.numParticles = d.numElements / 6;
		r.typeUboChanged = d.typeUboChanged;
		r.vertexByteOffset = d.vertexOffset;
		r.instanceUbo = d.instanceUbo;
	}
	return n;
}

void orc_effect_info(void *sys, uint32_t effectId, SprEffectInfo *out)
{
	RefSystem *rs = (RefSystem*)sys;
	const auto &e = rs->imp->effects[effectId];
	memset(out, 0, sizeof(*out));
	out->typeId = e.typeId;
	out->slotCount = e.particles.size * 4;
	out->aliveCount = e.gpuParticles.size / 4;
	out->freeCount = e.freeIndices.size;
	out->timeStep = e.timeStep;
	out->timeDelta = e.timeDelta;
	out->spawnTimer = e.spawnTimer;
	out->emitOnTimer = e.emitOnTimer;
	out->stopEmit = e.stopEmit;
	out->firstEmit = e.firstEmit;
	out->instantDelete = e.instantDelete;
	memcpy(&out->gpuBounds, &e.gpuBounds, sizeof(SprBounds3));
	out->lastUpdateTime = e.lastUpdateTime;
	out->simSteps = effectId < rs->simSteps.size() ? rs->simSteps[effectId] : 0;
}

uint32_t orc_read_free(void *sys, uint32_t effectId, uint32_t *out, uint32_t capacity)
{
	RefSystem *rs = (RefSystem*)sys;
	const auto &e = rs->imp->effects[effectId];
	uint32_t n = sf::min(e.freeIndices.size, capacity);
	memcpy(out, e.freeIndices.data, n * sizeof(uint32_t));
	return e.freeIndices.size;
}

uint32_t orc_read_stream(void *sys, uint32_t effectId, SprGpuParticle *out, uint32_t capacityVertices)
{
	RefSystem *rs = (RefSystem*)sys;
	const auto &e = rs->imp->effects[effectId];
	uint32_t n = sf::min(e.gpuParticles.size, capacityVertices);
	memcpy(out, e.gpuParticles.data, (size_t)n * sizeof(SprGpuParticle));
	return e.gpuParticles.size;
}

uint32_t orc_read_slots(void *sys, uint32_t effectId, SprGpuParticle *out, uint32_t capacitySlots)
{
	RefSystem *rs = (RefSystem*)sys;
	const auto &e = rs->imp->effects[effectId];
	uint32_t total = e.particles.size * 4;
	uint32_t n = sf::min(total, capacitySlots);
	for (uint32_t i = 0; i < n; i++) {
		const float *p4 = (const float*)&e.particles[i >> 2];
		uint32_t lane = i & 3;
		out[i].position.x = p4[0 + lane];
		out[i].position.y = p4[4 + lane];
		out[i].position.z = p4[8 + lane];
		out[i].velocity.x = p4[12 + lane];
		out[i].velocity.y = p4[16 + lane];
		out[i].velocity.z = p4[20 + lane];
		out[i].life = p4[24 + lane];
		out[i].seed = p4[28 + lane];
	}
	return total;
}

void orc_rng_state(void *sys, uint64_t outStateInc[2])
{
	RefSystem *rs = (RefSystem*)sys;
	outStateInc[0] = rs->imp->updateCtx.rng.state;
	outStateInc[1] = rs->imp->updateCtx.rng.inc;
}

void orc_type_lut(void *sys, uint32_t typeId, uint8_t out[SPR_SPLINE_LUT_BYTES])
{
	RefSystem *rs = (RefSystem*)sys;
	auto it = rs->luts.find(typeId);
	if (it != rs->luts.end()) memcpy(out, it->second.data(), SPR_SPLINE_LUT_BYTES);
	else memset(out, 0, SPR_SPLINE_LUT_BYTES);
}

void orc_type_ubo(void *sys, uint32_t typeId, SprVertexTypeUbo *out)
{
	RefSystem *rs = (RefSystem*)sys;
	memcpy(out, &rs->imp->types[typeId].typeUbo, sizeof(SprVertexTypeUbo));
}

#endif

// What the last orc_render submitted through sg_append_buffer / sg_apply_bindings, for the adapter-against-reference
// comparison of row N1. Vertex buffers are reported as ring indices 0..15 of the system (0xffffffff = none bound).
uint32_t orc_render_capture(void *sys, OrcDrawCapture *draws, uint32_t capDraws, OrcAppendCapture *appends, uint32_t capAppends, uint32_t *numAppends)
{
	RefSystem *rs = (RefSystem*)sys;
	auto ring = [&](uint32_t id) { return id >= rs->firstVertexBuffer && id < rs->firstVertexBuffer + 64 ? id - rs->firstVertexBuffer : 0xffffffffu; };
	uint32_t n = 0;
	for (; n < g_draws.size() && n < capDraws; n++) {
		const CapturedDraw &d = g_draws[n];
		draws[n] = { ring(d.vertexBuffer), d.vertexOffset, d.indexBuffer, d.vsImage0, d.vsImage1, d.fsImage0, d.numElements, 0 };
	}
	uint32_t m = 0;
	for (; m < g_appends.size() && m < capAppends; m++) appends[m] = { ring(g_appends[m].buffer), g_appends[m].offset, g_appends[m].numBytes, 0, g_appends[m].hash };
	if (numAppends) *numAppends = (uint32_t)g_appends.size();
	return (uint32_t)g_draws.size();
}

// The sphere areas currently registered with the (stub) area system: 5 floats per area {effect id, x, y, z, radius},
// ascending area id. Returns the number of areas.
uint32_t orc_area_spheres(void *sys, float *out, uint32_t capacity)
{
	RefSystem *rs = (RefSystem*)sys;
	StubAreaSystem *area = (StubAreaSystem*)rs->systems.area.ptr;
	uint32_t n = 0;
	for (const auto &kv : area->spheres) {
		if (n < capacity) { float *o = out + 5 * n; o[0] = (float)kv.second.userId; o[1] = kv.second.x; o[2] = kv.second.y; o[3] = kv.second.z; o[4] = kv.second.r; }
		n++;
	}
	return n;
}

// The 512 x 512 RGBA8 spline atlas (ParticleSystem.cpp:253-258, :421-449); returns 0 when nothing was ever copied into it.
int orc_read_atlas(void *sys, uint8_t *out)
{
	RefSystem *rs = (RefSystem*)sys;
	if (rs->atlas.empty()) { memset(out, 0, (size_t)kAtlasRes * kAtlasRes * 4); return 0; }
	memcpy(out, rs->atlas.data(), rs->atlas.size());
	return 1;
}

void orc_make_frustum(const SprMat44 *worldToClip, float clipNearW, SprFrustum *out)
{
	sf::Mat44 m;
	memcpy(m.v, worldToClip->v, sizeof(float) * 16);
	sf::Frustum f(m, clipNearW);
	for (int i = 0; i < 4; i++) {
		f.side[i].storeu(out->side[i]);
		f.caps[i].storeu(out->caps[i]);
	}
}

double orc_bench_steps(void *sys, const uint32_t *activeIds, uint32_t numActive,
	const SprFrameArgs *args, uint32_t steps, uint64_t *particleSteps)
{
	RefSystem *rs = (RefSystem*)sys;
	setActive(rs, activeIds, numActive);
	cl::FrameArgs fa = convertFrameArgs(*args);
	uint64_t total = 0;
	auto t0 = std::chrono::steady_clock::now();
	for (uint32_t s = 0; s < steps; s++) {
		fa.frameIndex++;
		fa.gameTime += (double)fa.dt;
		countSteps(rs, activeIds, numActive, fa.dt);
		rs->ps->updateParticles(rs->areas, fa);
#ifndef SPR_ADAPTER
		for (uint32_t i = 0; i < numActive; i++) total += rs->imp->effects[activeIds[i]].gpuParticles.size / 4;
#endif
	}
	auto t1 = std::chrono::steady_clock::now();
	if (particleSteps) *particleSteps += total;
	return std::chrono::duration<double>(t1 - t0).count();
}

}
