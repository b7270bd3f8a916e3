// descriptor_json.cpp -- SURVEY.md 8(f) row N3: effect descriptors from the reference's on-disk format.
//
// The reference stores prefabs as JSON written by sp::writeJson over its reflection tables
// (src/sp/Json.cpp; sv::ParticleSystemComponent: src/server/ServerStateReflection.cpp:321-363) and loads them with
// a lenient reader (loadConfig, src/server/ServerState.cpp:775-826: bare keys, // and /* */ comments, missing and
// trailing commas, control characters inside strings). This file reads that format into the flat
// SprParticleSystemComponent the C ABI takes -- host code, no GPU involved:
//   * a prefab            { name: .., components: [ { type: "ParticleSystem", ... }, ... ] }
//   * a tagged component  { type: "ParticleSystem", ... }
//   * a bare component    { timeStep: .., ... }
// with sp::readJson's rules (src/sp/Json.cpp:116-270): fields are looked up by name, absent ones keep the C++
// default, unknown keys are ignored, a value of the wrong JSON type or an unknown component type fails the load;
// numbers convert like (float)int64 / (float)double; a bool field also takes a number. sf::Symbol texture becomes
// the opaque 64-bit id of the POD: FNV-1a of the name (the name itself is kept and can be queried).

#include "particle_system.h"

#include <math.h>
#include <memory>
#include <stdlib.h>
#include <string.h>

namespace spr {

namespace {

struct JValue {
	enum Kind { Null, Bool, Int, Double, String, Array, Object } kind = Null;
	bool b = false;
	int64_t i = 0;
	double d = 0.0;
	std::string s;
	bool multiline = false;                                 // String whose first character is a raw line break (jsi_flag_multiline)
	std::vector<JValue> items;                              // Array
	std::vector<std::pair<std::string, JValue>> props;      // Object, in file order
	const JValue *get(const char *key) const {
		for (const auto &p : props) if (p.first == key) return &p.second;
		return nullptr;
	}
};

struct Parser {
	const char *p, *end;
	std::string error;
	int depth = 0;

	bool fail(const char *msg) { if (error.empty()) { error = msg; error += " at byte "; error += std::to_string((size_t)(p - begin)); } return false; }
	const char *begin;

	void skipWs() {
		for (;;) {
			while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) p++;
			if (p + 1 < end && p[0] == '/' && p[1] == '/') { while (p < end && *p != '\n') p++; continue; }
			if (p + 1 < end && p[0] == '/' && p[1] == '*') {
				p += 2;
				while (p + 1 < end && !(p[0] == '*' && p[1] == '/')) p++;
				if (p + 1 >= end) { p = end; return; }
				p += 2;
				continue;
			}
			return;
		}
	}

	static void appendUtf8(std::string &s, uint32_t c) {
		if (c < 0x80) s += (char)c;
		else if (c < 0x800) { s += (char)(0xc0 | (c >> 6)); s += (char)(0x80 | (c & 0x3f)); }
		else if (c < 0x10000) { s += (char)(0xe0 | (c >> 12)); s += (char)(0x80 | ((c >> 6) & 0x3f)); s += (char)(0x80 | (c & 0x3f)); }
		else { s += (char)(0xf0 | (c >> 18)); s += (char)(0x80 | ((c >> 12) & 0x3f)); s += (char)(0x80 | ((c >> 6) & 0x3f)); s += (char)(0x80 | (c & 0x3f)); }
	}

	// multiline (may be null): set when a RAW line break is the first thing in the string, alone or behind one raw carriage
	// return -- the reader's jsi_flag_multiline (src/ext/json_input.cpp.h:689-693); an escaped \n does not count
	bool parseString(std::string &out, bool *multiline = nullptr) { // p is after the opening quote
		while (p < end) {
			char c = *p++;
			if (c == '"') return true;
			if (c == '\\') {
				if (p >= end) break;
				char e = *p++;
				switch (e) {
				case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
				case 'b': out += '\b'; break; case 'f': out += '\f'; break;
				case '"': out += '"'; break; case '\\': out += '\\'; break; case '/': out += '/'; break;
				case 'u': {
					if (end - p < 4) return fail("truncated \\u escape");
					uint32_t c32 = 0;
					for (int k = 0; k < 4; k++) {
						char h = *p++;
						c32 = c32 * 16 + (uint32_t)(h >= '0' && h <= '9' ? h - '0' : h >= 'a' && h <= 'f' ? h - 'a' + 10 : h >= 'A' && h <= 'F' ? h - 'A' + 10 : 0x10000);
					}
					if (c32 >= 0x10000 * 16) return fail("bad \\u escape");
					appendUtf8(out, c32);
				} break;
				default: return fail("bad escape");
				}
			} else {
				if (c == '\n' && multiline && (out.empty() || out == "\r")) *multiline = true;
				out += c; // control characters are allowed inside strings (allow_control_in_string)
			}
		}
		return fail("unclosed string");
	}

	bool parseValue(JValue &v) {
		if (++depth > 64) return fail("nesting too deep");
		skipWs();
		if (p >= end) return fail("unexpected end");
		bool ok = true;
		char c = *p;
		if (c == '{') {
			p++;
			v.kind = JValue::Object;
			for (;;) {
				skipWs();
				if (p >= end) { ok = fail("unclosed object"); break; }
				if (*p == '}') { p++; break; }          // also after a trailing comma
				if (*p == ',') { p++; continue; }        // separators are optional (allow_missing_comma)
				std::string key;
				if (*p == '"') { p++; if (!parseString(key)) { ok = false; break; } }
				else {
					while (p < end && ((*p >= 'A' && *p <= 'Z') || (*p >= 'a' && *p <= 'z') || (*p >= '0' && *p <= '9') || *p == '_' || *p == '$')) key += *p++;
					if (key.empty()) { ok = fail("expected a key or '}'"); break; }
				}
				skipWs();
				if (p >= end || *p != ':') { ok = fail("expected ':' after key"); break; }
				p++;
				v.props.emplace_back(std::move(key), JValue());
				if (!parseValue(v.props.back().second)) { ok = false; break; }
			}
		} else if (c == '[') {
			p++;
			v.kind = JValue::Array;
			for (;;) {
				skipWs();
				if (p >= end) { ok = fail("unclosed array"); break; }
				if (*p == ']') { p++; break; }
				if (*p == ',') { p++; continue; }
				v.items.emplace_back();
				if (!parseValue(v.items.back())) { ok = false; break; }
			}
		} else if (c == '"') {
			p++;
			v.kind = JValue::String;
			ok = parseString(v.s, &v.multiline);
		} else if (c == '-' || (c >= '0' && c <= '9')) {
			const char *q = p;
			if (*q == '-') q++;
			while (q < end && *q >= '0' && *q <= '9') q++;
			bool isInt = true;
			if (q < end && *q == '.') { isInt = false; q++; while (q < end && *q >= '0' && *q <= '9') q++; }
			if (q < end && (*q == 'e' || *q == 'E')) { isInt = false; q++; if (q < end && (*q == '+' || *q == '-')) q++; while (q < end && *q >= '0' && *q <= '9') q++; }
			std::string num(p, q);
			if (isInt && num.size() <= 18) { v.kind = JValue::Int; v.i = strtoll(num.c_str(), nullptr, 10); }
			else { v.kind = JValue::Double; v.d = strtod(num.c_str(), nullptr); }
			p = q;
		} else if (end - p >= 4 && !memcmp(p, "true", 4)) { v.kind = JValue::Bool; v.b = true; p += 4; }
		else if (end - p >= 5 && !memcmp(p, "false", 5)) { v.kind = JValue::Bool; v.b = false; p += 5; }
		else if (end - p >= 4 && !memcmp(p, "null", 4)) { v.kind = JValue::Null; p += 4; }
		else ok = fail("invalid character");
		depth--;
		return ok;
	}
};

// ---- sp::readJson over the particle component's reflection table

struct Reader {
	std::string error;
	bool fail(const std::string &what) { if (error.empty()) error = what; return false; }

	// IsPrimitive branch (Json.cpp:226-262)
	template <typename T> bool number(const JValue *v, const char *name, T &dst) {
		if (!v) return true;
		if (v->kind == JValue::Int) dst = (T)v->i;
		else if (v->kind == JValue::Double) dst = (T)v->d;
		else return fail(std::string("field '") + name + "' expects a number");
		return true;
	}
	bool boolean(const JValue *v, const char *name, uint8_t &dst) {
		if (!v) return true;
		if (v->kind == JValue::Bool) dst = v->b ? 1 : 0;
		else if (v->kind == JValue::Int) dst = v->i != 0;
		else if (v->kind == JValue::Double) dst = v->d != 0.0;
		else return fail(std::string("field '") + name + "' expects a boolean");
		return true;
	}
	bool object(const JValue *v, const char *name) {
		if (v->kind != JValue::Object) return fail(std::string("field '") + name + "' expects an object");
		return true;
	}
	bool vec2(const JValue *v, const char *name, SprVec2 &dst) {
		if (!v) return true;
		return object(v, name) && number(v->get("x"), name, dst.x) && number(v->get("y"), name, dst.y);
	}
	bool vec3(const JValue *v, const char *name, SprVec3 &dst) {
		if (!v) return true;
		return object(v, name) && number(v->get("x"), name, dst.x) && number(v->get("y"), name, dst.y) && number(v->get("z"), name, dst.z);
	}
	bool randomVec3(const JValue *v, const char *name, SprRandomVec3 &dst) { // sv::RandomVec3 / RandomSphere, ServerState.h:161-178
		if (!v) return true;
		if (!object(v, name)) return false;
		if (!vec3(v->get("offset"), "offset", dst.offset) || !vec3(v->get("boxExtent"), "boxExtent", dst.boxExtent)) return false;
		if (const JValue *s = v->get("sphere")) {
			if (!object(s, "sphere")) return false;
			if (!number(s->get("minTheta"), "minTheta", dst.sphere.minTheta) || !number(s->get("maxTheta"), "maxTheta", dst.sphere.maxTheta)
				|| !number(s->get("minPhi"), "minPhi", dst.sphere.minPhi) || !number(s->get("maxPhi"), "maxPhi", dst.sphere.maxPhi)
				|| !number(s->get("minRadius"), "minRadius", dst.sphere.minRadius) || !number(s->get("maxRadius"), "maxRadius", dst.sphere.maxRadius)
				|| !vec3(s->get("scale"), "scale", dst.sphere.scale)) return false;
		}
		return vec3(v->get("rotation"), "rotation", dst.rotation);
	}
};

} // namespace

struct DescriptorSet {
	struct Item {
		SprParticleSystemComponent comp;
		std::string texture;
		std::vector<SprVec2> splines[4];
		std::vector<SprGradientPoint> gradient;
		std::vector<SprGravityPoint> gravityPoints;
	};
	std::vector<std::unique_ptr<Item>> items; // stable addresses: the descriptor pointer is the effect type key
};

static uint64_t fnv1a64(const std::string &s)
{
	uint64_t h = 0xcbf29ce484222325ull;
	for (unsigned char c : s) { h ^= c; h *= 0x100000001b3ull; }
	return h;
}

// What sp::readJson stores for a string value flagged multiline (src/sp/Json.cpp:134-166): the line break that opens the
// string is dropped, the indentation of the first line (its count of spaces and its count of tabs) is the indentation that
// is taken off every following line -- as long as either count is unmet, blanks of both kinds are eaten --, lines are
// joined with '\n', and a last line that holds nothing but indentation (the one in front of the closing quote) is dropped.
static std::string dedentMultiline(const std::string &s)
{
	const char *p = s.c_str(); // NUL-terminated: an embedded NUL ends the value, as it does in the reference's C string
	if (*p == '\r') p++;
	if (*p == '\n') p++;
	uint32_t wantSpaces = 0, wantTabs = 0;
	for (;; p++) {
		if (*p == ' ') wantSpaces++;
		else if (*p == '\t') wantTabs++;
		else break;
	}
	std::string out;
	auto copyLine = [&] { while (*p != '\r' && *p != '\n' && *p != '\0') out += *p++; };
	copyLine();
	while (*p != '\0') {
		if (*p == '\r') p++;
		if (*p == '\n') p++;
		uint32_t spaces = 0, tabs = 0;
		for (; spaces < wantSpaces || tabs < wantTabs; p++) {
			if (*p == ' ') spaces++;
			else if (*p == '\t') tabs++;
			else break;
		}
		if (*p == '\0') break;
		out += '\n';
		copyLine();
	}
	return out;
}

static bool readComponent(Reader &R, const JValue &o, DescriptorSet::Item &it)
{
	SprParticleSystemComponent &c = it.comp;
	sprComponentDefaults(&c); // the C++ defaults of ServerState.h:187-239: absent fields keep them
	if (const JValue *t = o.get("texture")) {
		if (t->kind == JValue::String) it.texture = t->multiline ? dedentMultiline(t->s) : t->s;
		else if (t->kind != JValue::Null) return R.fail("field 'texture' expects a string");
	}
	c.texture = it.texture.empty() ? 0 : fnv1a64(it.texture);
	if (const JValue *f = o.get("frameCount")) {
		if (!R.object(f, "frameCount") || !R.number(f->get("x"), "frameCount", c.frameCount[0]) || !R.number(f->get("y"), "frameCount", c.frameCount[1])) return false;
	}
#define NUM(field) if (!R.number(o.get(#field), #field, c.field)) return false
#define BOOL(field) if (!R.boolean(o.get(#field), #field, c.field)) return false
	NUM(timeStep); NUM(prewarmTime); NUM(updateRadius); NUM(cullPadding); BOOL(updateOutOfCamera); NUM(renderOrder);
	NUM(spawnTime); NUM(spawnTimeVariance); NUM(burstAmount); NUM(burstAmountVariance); NUM(emitterOnTime);
	BOOL(localSpace); BOOL(instantDelete);
	if (!R.randomVec3(o.get("emitPosition"), "emitPosition", c.emitPosition) || !R.randomVec3(o.get("emitVelocity"), "emitVelocity", c.emitVelocity)) return false;
	if (!R.vec3(o.get("emitVelocityAttractorOffset"), "emitVelocityAttractorOffset", c.emitVelocityAttractorOffset)) return false;
	NUM(emitVelocityAttractorStrength);
	if (const JValue *a = o.get("gravityPoints")) {
		if (a->kind != JValue::Array) return R.fail("field 'gravityPoints' expects an array");
		for (const JValue &e : a->items) {
			SprGravityPoint gp = { { 0, 0, 0 }, 0.5f, 1.0f }; // ServerState.h:180-185
			if (!R.object(&e, "gravityPoints[]") || !R.vec3(e.get("position"), "position", gp.position) || !R.number(e.get("radius"), "radius", gp.radius)
				|| !R.number(e.get("strength"), "strength", gp.strength)) return false;
			it.gravityPoints.push_back(gp);
		}
	}
	NUM(drag);
	if (!R.vec3(o.get("gravity"), "gravity", c.gravity)) return false;
	NUM(size); NUM(sizeVariance); NUM(lifeTime); NUM(lifeTimeVariance);
	const char *splineNames[4] = { "scaleSpline", "alphaSpline", "additiveSpline", "erosionSpline" };
	for (int k = 0; k < 4; k++) {
		const JValue *sp = o.get(splineNames[k]);
		if (!sp) continue;
		if (!R.object(sp, splineNames[k])) return false;
		if (const JValue *pts = sp->get("points")) {
			if (pts->kind != JValue::Array) return R.fail("spline 'points' expects an array");
			for (const JValue &e : pts->items) {
				SprVec2 p = { 0, 0 };
				if (!R.vec2(&e, "points[]", p)) return false;
				it.splines[k].push_back(p);
			}
		}
	}
	if (const JValue *g = o.get("gradient")) {
		if (!R.object(g, "gradient") || !R.vec3(g->get("defaultColor"), "defaultColor", c.gradient.defaultColor)) return false;
		if (const JValue *pts = g->get("points")) {
			if (pts->kind != JValue::Array) return R.fail("gradient 'points' expects an array");
			for (const JValue &e : pts->items) {
				SprGradientPoint p = { 0.0f, { 0, 0, 0 } }; // ServerState.h:149-153
				if (!R.object(&e, "points[]") || !R.number(e.get("t"), "t", p.t) || !R.vec3(e.get("color"), "color", p.color)) return false;
				it.gradient.push_back(p);
			}
		}
	}
	NUM(frameRate); BOOL(relativeFrameRate); BOOL(randomStartFrame);
	NUM(rotation); NUM(rotationVariance); NUM(spin); NUM(spinVariance);
#undef NUM
#undef BOOL
	SprBSpline2 *splines[4] = { &c.scaleSpline, &c.alphaSpline, &c.additiveSpline, &c.erosionSpline };
	for (int k = 0; k < 4; k++) { splines[k]->points = it.splines[k].empty() ? nullptr : it.splines[k].data(); splines[k]->numPoints = (uint32_t)it.splines[k].size(); }
	c.gradient.points = it.gradient.empty() ? nullptr : it.gradient.data();
	c.gradient.numPoints = (uint32_t)it.gradient.size();
	c.gravityPoints = it.gravityPoints.empty() ? nullptr : it.gravityPoints.data();
	c.numGravityPoints = (uint32_t)it.gravityPoints.size();
	return true;
}

// the polymorphic sv::Component types sp::readJson knows (ServerStateReflection.cpp:11-47): any other tag fails the load
static const char *const kComponentTypes[] = {
	"DynamicModel", "TileModel", "PointLight", "ParticleSystem", "Character", "CharacterModel", "TapArea", "Door", "Chest", "BlobShadow",
	"Card", "CardAttach", "CardStatus", "CardMelee", "CardKey", "Projectile", "DamageOnTurnStart", "CastOnTurnStart", "CastOnReceiveDamage",
	"CastOnDealDamage", "ResistDamage", "IncreaseDamage", "CardCast", "CardCastMelee", "Spell", "SpellDamage", "SpellHeal", "SpellStatus",
	"Status", "StatusChangeTeam", "CharacterTemplate", "TileArea" };

// a polymorphic Box<Component> (Json.cpp:178-194): returns false on error; appends when it is a particle system
static bool readTagged(Reader &R, const JValue &v, DescriptorSet &set)
{
	if (v.kind == JValue::Null) return true;
	if (v.kind != JValue::Object) return R.fail("a component expects an object");
	const JValue *tag = v.get("type");
	if (!tag || tag->kind != JValue::String) return R.fail("a component needs a string 'type'");
	bool known = false;
	for (const char *n : kComponentTypes) known |= tag->s == n;
	if (!known) return R.fail("unknown component type '" + tag->s + "'");
	if (tag->s != "ParticleSystem") return true; // another system's component: not ours to read
	std::unique_ptr<DescriptorSet::Item> it(new DescriptorSet::Item());
	if (!readComponent(R, v, *it)) return false;
	set.items.push_back(std::move(it));
	return true;
}

DescriptorSet *descriptorsFromJson(const char *json, size_t length, std::string &error)
{
	Parser P;
	P.p = P.begin = json; P.end = json + length;
	JValue root;
	if (!P.parseValue(root)) { error = "JSON: " + P.error; return nullptr; }
	P.skipWs();
	if (root.kind != JValue::Object) { error = "JSON: the document is not an object"; return nullptr; }
	std::unique_ptr<DescriptorSet> set(new DescriptorSet());
	Reader R;
	bool ok = true;
	if (const JValue *comps = root.get("components")) {          // sv::Prefab, ServerState.h:568-572
		if (comps->kind != JValue::Array) ok = R.fail("'components' expects an array");
		else for (const JValue &c : comps->items) if (!(ok = readTagged(R, c, *set))) break;
	} else if (root.get("type")) {
		ok = readTagged(R, root, *set);
	} else {
		std::unique_ptr<DescriptorSet::Item> it(new DescriptorSet::Item());
		ok = readComponent(R, root, *it);
		if (ok) set->items.push_back(std::move(it));
	}
	if (!ok) { error = R.error; return nullptr; }
	return set.release();
}

uint32_t descriptorCount(const DescriptorSet *s) { return (uint32_t)s->items.size(); }
const SprParticleSystemComponent *descriptorGet(const DescriptorSet *s, uint32_t i) { return i < s->items.size() ? &s->items[i]->comp : nullptr; }
const char *descriptorTexture(const DescriptorSet *s, uint32_t i) { return i < s->items.size() ? s->items[i]->texture.c_str() : nullptr; }
void descriptorsFree(DescriptorSet *s) { delete s; }

} // namespace spr
