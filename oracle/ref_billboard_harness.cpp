// ref_billboard_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Headless harness around the UNMODIFIED reference cl::BillboardSystem: this TU includes
// /root/reference/src/client/BillboardSystem.cpp (found through the include path set up by
// oracle/build_ref.py; no reference source is copied into this repository) and supplies link
// stubs for the renderer / asset services that file references. The sg_* stubs capture what
// renderMain submits: the vertex buffer contents (sg_update_buffer, :206), the uniform block
// (:208-210) and one (atlas, count, offset) per sg_draw (:213-225). Oracle for SURVEY.md 8(f) row N4.

#include "client/BillboardSystem.h"
#include "sp/Sprite.h"
#include "sp/Renderer.h"
#include "game/shader/GameShaders.h"
#include "game/shader/Billboard.h"
#include "spear_particles.h"

#include "client/BillboardSystem.cpp"

#include <unordered_map>
#include <vector>
#include <memory>
#include <string.h>

GameShaders gameShaders;
void GameShaders::load() { }

namespace sp {

sg_buffer g_hackSharedQuadIndexBuffer;

Buffer::Buffer() { }
Buffer::~Buffer() { }
void Buffer::reset() { }
void Buffer::initDynamicVertex(const char *, size_t) { buffer.id = 7; }

Pipeline::Pipeline() { }
Pipeline::~Pipeline() { }
bool Pipeline::bind() { return true; }
void Pipeline::reset() { }

AssetProps::~AssetProps() { }
uint32_t SpriteProps::hash() const { return 0; }
bool SpriteProps::equal(const AssetProps &) const { return true; }
void SpriteProps::copyTo(AssetProps *) const { }

Asset::Asset() { }
Asset::~Asset() { }
bool Asset::isLoaded() const { return true; }   // every sprite the harness hands over is "loaded" (:82, :100)

AssetType Sprite::SelfType;

}

namespace cl {
System::~System() { }
}

namespace {

struct HSprite final : sp::Sprite
{
	void assetStartLoading() override { }
	void assetUnload() override { }
};

struct Captured
{
	std::vector<uint8_t> vertices;
	float worldToClip[16];
	struct Draw { uint32_t image; uint32_t numElements; uint32_t offset; };
	std::vector<Draw> draws;
	uint32_t curImage = 0, curOffset = 0;
};
Captured g_cap;

struct Harness
{
	sf::Box<cl::BillboardSystem> system;
	std::unordered_map<uint64_t, std::unique_ptr<sp::Atlas>> atlases;   // atlas id -> sp::Atlas (image.id = running number)
	std::unordered_map<uint32_t, uint64_t> imageToAtlas;
	std::vector<std::unique_ptr<HSprite>> sprites;                       // sprites of the current frame
};

}

extern "C" {

void sg_update_buffer(sg_buffer, const void *data, int numBytes)
{
	g_cap.vertices.assign((const uint8_t*)data, (const uint8_t*)data + numBytes);
}
void sg_apply_uniforms(sg_shader_stage, int slot, const void *data, int numBytes)
{
	if (slot == SLOT_Billboard_Vertex && numBytes == (int)sizeof(g_cap.worldToClip)) memcpy(g_cap.worldToClip, data, sizeof(g_cap.worldToClip));
}
void sg_apply_bindings(const sg_bindings *b)
{
	g_cap.curImage = b->fs_images[SLOT_Billboard_u_Atlas].id;
	g_cap.curOffset = (uint32_t)b->vertex_buffer_offsets[0];
}
void sg_draw(int, int numElements, int)
{
	g_cap.draws.push_back({ g_cap.curImage, (uint32_t)numElements, g_cap.curOffset });
}

#define ORC_EXPORT __attribute__((visibility("default")))

ORC_EXPORT void *orc_bb_create()
{
	Harness *h = new Harness();
	h->system = cl::BillboardSystem::create();
	return h;
}

ORC_EXPORT void orc_bb_destroy(void *p) { delete (Harness*)p; }

// cl::BillboardSystem::addBillboard(const Billboard&) for each entry (atlas 0 = null sprite)
ORC_EXPORT void orc_bb_add(void *p, const SprBillboard *bb, uint32_t count)
{
	Harness *h = (Harness*)p;
	for (uint32_t i = 0; i < count; i++) {
		const SprBillboard &s = bb[i];
		cl::Billboard b;
		if (s.sprite.atlas) {
			auto &atlas = h->atlases[s.sprite.atlas];
			if (!atlas) {
				atlas.reset(new sp::Atlas());
				atlas->width = 1024; atlas->height = 512;
				atlas->image.id = (uint32_t)h->atlases.size() + 100;
				h->imageToAtlas[atlas->image.id] = s.sprite.atlas;
			}
			h->sprites.emplace_back(new HSprite());
			HSprite *sp = h->sprites.back().get();
			sp->atlas = atlas.get();
			sp->minVert = sf::Vec2(s.sprite.minVert.x, s.sprite.minVert.y);
			sp->maxVert = sf::Vec2(s.sprite.maxVert.x, s.sprite.maxVert.y);
			sp->vertUvScale = sf::Vec2(s.sprite.vertUvScale.x, s.sprite.vertUvScale.y);
			sp->vertUvBias = sf::Vec2(s.sprite.vertUvBias.x, s.sprite.vertUvBias.y);
			b.sprite = sp;
		}
		memcpy(b.transform.v, s.transform, sizeof(float) * 12);
		b.color = sf::Vec4(s.color.x, s.color.y, s.color.z, s.color.w);
		b.depth = s.depth;
		b.anchor = sf::Vec2(s.anchor.x, s.anchor.y);
		b.cropMin = sf::Vec2(s.cropMin.x, s.cropMin.y);
		b.cropMax = sf::Vec2(s.cropMax.x, s.cropMax.y);
		h->system->addBillboard(b);
	}
}

// cl::BillboardSystem::renderMain; returns the number of quads in the submitted vertex buffer
ORC_EXPORT uint32_t orc_bb_render(void *p, const SprRenderArgs *args, SprBillboardVertex *outVerts, uint32_t capQuads,
	SprBillboardDraw *outDraws, uint32_t capDraws, uint32_t *outNumDraws, float outWorldToClip[16])
{
	Harness *h = (Harness*)p;
	g_cap = Captured();
	cl::RenderArgs ra;
	memcpy(ra.worldToClip.v, args->worldToClip.v, sizeof(float) * 16);
	cl::VisibleAreas areas;
	h->system->renderMain(areas, ra);
	uint32_t quads = (uint32_t)(g_cap.vertices.size() / (4 * sizeof(SprBillboardVertex)));
	if (outVerts) memcpy(outVerts, g_cap.vertices.data(), (size_t)(quads < capQuads ? quads : capQuads) * 4 * sizeof(SprBillboardVertex));
	uint32_t nd = 0;
	for (const Captured::Draw &d : g_cap.draws) {
		if (nd < capDraws && outDraws) {
			outDraws[nd].atlas = h->imageToAtlas[d.image];
			outDraws[nd].count = d.numElements / 6;
			outDraws[nd].firstQuad = d.offset / (4 * (uint32_t)sizeof(SprBillboardVertex));
		}
		nd++;
	}
	if (outNumDraws) *outNumDraws = nd;
	if (outWorldToClip) memcpy(outWorldToClip, g_cap.worldToClip, sizeof(g_cap.worldToClip));
	h->sprites.clear();
	return quads;
}

}

static_assert(sizeof(SprBillboardVertex) == 24, "BillboardVertex is 24 bytes (src/client/BillboardSystem.cpp:53-58)");
