// ref_area_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Oracle of SURVEY.md 8(f) row N5: the reference's own cl::AreaSystem (src/client/AreaSystem.cpp, compiled where it lies
// under /root/reference by oracle/build_ref.py) holding the sphere areas the particle system registers
// (ParticleSystem.cpp:762-763, :782-783, :801), queried with sf::Frustum (src/sf/Frustum.h).

#include "client/AreaSystem.h"
#include "sf/Frustum.h"
#include "spear_particles.h"

#include <string.h>

namespace cl {
System::~System() { }
}

namespace {
struct Harness { sf::Box<cl::AreaSystem> area; };
}

#define ORC_EXPORT extern "C" __attribute__((visibility("default")))

ORC_EXPORT void *orc_area_create() { Harness *h = new Harness(); h->area = cl::AreaSystem::create(); return h; }
ORC_EXPORT void orc_area_destroy(void *p) { delete (Harness*)p; }

// addSphereArea(AreaGroup::ParticleEffect, effectId, { position, updateRadius }, Activate | Visibility) -> areaId (:762-763);
// `group` lets a test register areas of other systems next to the particle ones
ORC_EXPORT uint32_t orc_area_add_sphere(void *p, uint32_t group, uint32_t userId, const float origin[3], float radius)
{
	sf::Sphere s = { sf::Vec3(origin[0], origin[1], origin[2]), radius };
	return ((Harness*)p)->area->addSphereArea((cl::AreaGroup)group, userId, s, cl::Area::Activate | cl::Area::Visibility);
}
ORC_EXPORT void orc_area_update_sphere(void *p, uint32_t areaId, const float origin[3], float radius)
{
	sf::Sphere s = { sf::Vec3(origin[0], origin[1], origin[2]), radius };
	((Harness*)p)->area->updateSphereArea(areaId, s);
}
ORC_EXPORT void orc_area_remove_sphere(void *p, uint32_t areaId) { ((Harness*)p)->area->removeSphereArea(areaId); }
ORC_EXPORT void orc_area_optimize(void *p) { ((Harness*)p)->area->optimize(); }

// queryFrustum(areas, Area::Visibility, frustum), userIds of AreaGroup::ParticleEffect in the reference's own order
ORC_EXPORT uint32_t orc_area_query(void *p, const SprFrustum *f, uint32_t *out, uint32_t capacity)
{
	sf::Frustum fr;
	for (int k = 0; k < 4; k++) { fr.side[k] = sf::Float4::loadu(f->side[k]); fr.caps[k] = sf::Float4::loadu(f->caps[k]); }
	sf::Array<cl::Area> areas;
	((Harness*)p)->area->queryFrustum(areas, cl::Area::Visibility, fr);
	uint32_t n = 0;
	for (const cl::Area &a : areas) {
		if (a.group != cl::AreaGroup::ParticleEffect) continue;
		if (n < capacity) out[n] = a.userId;
		n++;
	}
	return n;
}

// the enum values tests need (src/client/System.h:157-168)
ORC_EXPORT uint32_t orc_area_particle_group() { return (uint32_t)cl::AreaGroup::ParticleEffect; }
ORC_EXPORT uint32_t orc_area_other_group() { return (uint32_t)cl::AreaGroup::PointLight; }
