// ref_json_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Oracle of SURVEY.md 8(f) row N3 (effect-descriptor loader): the reference's own on-disk path for a
// sv::ParticleSystemComponent -- the lenient JSON reader with the dialect of loadConfig
// (src/server/ServerState.cpp:786-792), sp::readJson over the reflection tables of
// src/server/ServerStateReflection.cpp (:321-363 for the particle component), and sp::writeJson for producing
// fixtures in the reference's own canonical form. The reference units are compiled where they lie under
// /root/reference by oracle/build_ref.py; nothing is copied.
//
// The component is handed back as a flat vector of doubles in declaration order (ServerState.h:187-239; arrays as
// count + elements; float -> double is exact), which the product's loader is flattened to as well.

#include "server/ServerState.h"
#include "sp/Json.h"
#include "sf/Reflection.h"
#include "spear_particles.h"

#include "server/ServerStateReflection.h"

#include <string.h>
#include <vector>

// Link stub: editor metadata of the reflected fields (descriptions, asset / colour hints; the definition lives in
// ServerState.cpp next to the game logic). Nothing on the JSON path reads it.
namespace sv {
ReflectionInfo &addTypeReflectionInfo(sf::Type *, const sf::String &)
{
	static ReflectionInfo scratch;
	scratch = ReflectionInfo();
	return scratch;
}
}

namespace {

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
	for (uint32_t i = 0; i < s.numPoints; i++) d.points.push(sf::Vec2(s.points[i].x, s.points[i].y));
}

sf::Box<sv::ParticleSystemComponent> convertComponent(const SprParticleSystemComponent &s, const char *texture)
{
	sf::Box<sv::ParticleSystemComponent> box = sf::box<sv::ParticleSystemComponent>();
	sv::ParticleSystemComponent &d = *box;
	if (texture) d.texture = sf::Symbol(texture);
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

uint64_t fnv1a64(const char *s, size_t n)
{
	uint64_t h = 0xcbf29ce484222325ull;
	for (size_t i = 0; i < n; i++) { h ^= (uint8_t)s[i]; h *= 0x100000001b3ull; }
	return h;
}

void put3(std::vector<double> &o, const sf::Vec3 &v) { o.push_back(v.x); o.push_back(v.y); o.push_back(v.z); }

void putRandomVec3(std::vector<double> &o, const sv::RandomVec3 &r)
{
	put3(o, r.offset); put3(o, r.boxExtent);
	o.push_back(r.sphere.minTheta); o.push_back(r.sphere.maxTheta); o.push_back(r.sphere.minPhi); o.push_back(r.sphere.maxPhi);
	o.push_back(r.sphere.minRadius); o.push_back(r.sphere.maxRadius); put3(o, r.sphere.scale);
	put3(o, r.rotation);
}

void putSpline(std::vector<double> &o, const sv::BSpline2 &s)
{
	o.push_back((double)s.points.size);
	for (const sf::Vec2 &p : s.points) { o.push_back(p.x); o.push_back(p.y); }
}

// declaration order of ServerState.h:187-239
void flatten(std::vector<double> &o, const sv::ParticleSystemComponent &c)
{
	uint64_t tex = c.texture.size() ? fnv1a64(c.texture.data, c.texture.size()) : 0;
	o.push_back((double)(uint32_t)(tex >> 32)); o.push_back((double)(uint32_t)tex);
	o.push_back(c.frameCount.x); o.push_back(c.frameCount.y);
	o.push_back(c.timeStep); o.push_back(c.prewarmTime); o.push_back(c.updateRadius); o.push_back(c.cullPadding);
	o.push_back(c.updateOutOfCamera ? 1 : 0); o.push_back(c.renderOrder);
	o.push_back(c.spawnTime); o.push_back(c.spawnTimeVariance); o.push_back(c.burstAmount); o.push_back(c.burstAmountVariance);
	o.push_back(c.emitterOnTime); o.push_back(c.localSpace ? 1 : 0); o.push_back(c.instantDelete ? 1 : 0);
	putRandomVec3(o, c.emitPosition); putRandomVec3(o, c.emitVelocity);
	put3(o, c.emitVelocityAttractorOffset); o.push_back(c.emitVelocityAttractorStrength);
	o.push_back((double)c.gravityPoints.size);
	for (const sv::GravityPoint &p : c.gravityPoints) { put3(o, p.position); o.push_back(p.radius); o.push_back(p.strength); }
	o.push_back(c.drag); put3(o, c.gravity);
	o.push_back(c.size); o.push_back(c.sizeVariance); o.push_back(c.lifeTime); o.push_back(c.lifeTimeVariance);
	putSpline(o, c.scaleSpline); putSpline(o, c.alphaSpline); putSpline(o, c.additiveSpline); putSpline(o, c.erosionSpline);
	put3(o, c.gradient.defaultColor);
	o.push_back((double)c.gradient.points.size);
	for (const sv::GradientPoint &p : c.gradient.points) { o.push_back(p.t); put3(o, p.color); }
	o.push_back(c.frameRate); o.push_back(c.relativeFrameRate ? 1 : 0); o.push_back(c.randomStartFrame ? 1 : 0);
	o.push_back(c.rotation); o.push_back(c.rotationVariance); o.push_back(c.spin); o.push_back(c.spinVariance);
}

}

#define ORC_EXPORT extern "C" __attribute__((visibility("default")))

// Parses `json` the way loadConfig does (ServerState.cpp:786-796) and reads it through the reflection tables:
// a prefab ({ components: [ { type: "ParticleSystem", ... }, ... ] }), a tagged component ({ type: "ParticleSystem", ... })
// or a bare sv::ParticleSystemComponent object. Every particle component found is appended to `out` as
// [length, fields...]. Returns the number of components, -1 on a parse / read error.
ORC_EXPORT int orc_json_read(const char *json, size_t length, double *out, size_t capacity, size_t *outLength)
{
	jsi_args args = { };
	args.dialect.allow_bare_keys = true;
	args.dialect.allow_comments = true;
	args.dialect.allow_control_in_string = true;
	args.dialect.allow_missing_comma = true;
	args.dialect.allow_trailing_comma = true;
	jsi_value *value = jsi_parse_memory(json, length, &args);
	if (!value) return -1;
	std::vector<double> flat;
	int count = 0;
	bool ok = true;
	if (value->type != jsi_type_object) ok = false;
	else if (jsi_get(value->object, "components")) {
		sv::Prefab prefab;
		ok = sp::readJson(value, prefab);
		if (ok) for (sf::Box<sv::Component> &c : prefab.components) {
			if (!c || c->type != sv::Component::ParticleSystem) continue;
			std::vector<double> one;
			flatten(one, *(sv::ParticleSystemComponent*)c.ptr);
			flat.push_back((double)one.size()); flat.insert(flat.end(), one.begin(), one.end());
			count++;
		}
	} else if (jsi_get(value->object, "type")) {
		sf::Box<sv::Component> c;
		ok = sp::readJson(value, c);
		if (ok && c && c->type == sv::Component::ParticleSystem) {
			std::vector<double> one;
			flatten(one, *(sv::ParticleSystemComponent*)c.ptr);
			flat.push_back((double)one.size()); flat.insert(flat.end(), one.begin(), one.end());
			count++;
		}
	} else {
		sv::ParticleSystemComponent c;
		ok = sp::readJson(value, c);
		if (ok) {
			std::vector<double> one;
			flatten(one, c);
			flat.push_back((double)one.size()); flat.insert(flat.end(), one.begin(), one.end());
			count++;
		}
	}
	jsi_free(value);
	if (!ok) return -1;
	if (outLength) *outLength = flat.size();
	if (out) memcpy(out, flat.data(), sizeof(double) * (flat.size() < capacity ? flat.size() : capacity));
	return count;
}

// sp::writeJson of a prefab holding the given components (texture names optional): the reference's own canonical text.
ORC_EXPORT size_t orc_json_write_prefab(const SprParticleSystemComponent *const *comps, const char *const *textures, uint32_t count, int pretty,
	char *out, size_t capacity)
{
	sv::Prefab prefab;
	prefab.name = sf::Symbol("oracle/prefab");
	for (uint32_t i = 0; i < count; i++) prefab.components.push(convertComponent(*comps[i], textures ? textures[i] : nullptr));
	sf::Array<char> text;
	jso_stream s;
	sp::jsoInitArray(&s, text);
	s.pretty = pretty != 0;
	sp::writeJson(s, prefab);
	jso_close(&s);
	if (out && capacity) { size_t n = text.size < capacity - 1 ? text.size : capacity - 1; memcpy(out, text.data, n); out[n] = 0; }
	return text.size;
}

// the same flattening applied to a POD descriptor (what the product's loader output is compared through)
ORC_EXPORT size_t orc_json_flatten_pod(const SprParticleSystemComponent *c, const char *texture, double *out, size_t capacity)
{
	std::vector<double> one;
	flatten(one, *convertComponent(*c, texture));
	if (out) memcpy(out, one.data(), sizeof(double) * (one.size() < capacity ? one.size() : capacity));
	return one.size();
}
