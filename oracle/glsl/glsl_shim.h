// glsl_shim.h -- TEST INFRASTRUCTURE ONLY: just enough GLSL 3.30 semantics in C++ to EXECUTE the reference's own
// vertex shader text on the CPU (row A12 / N2 of SURVEY.md section 8). There is no GL driver in this environment
// (profiles/r02_probe_gl.md), so instead of a restatement of the shader, the shader source the reference SHIPS --
// Particle_vs_source_glsl330 in /root/reference/src/game/shader/Particle.h:239, the GLSL its GL backend hands to
// the driver, generated from Particle.glsl:73-140 -- is decoded at build time (oracle/glsl/build_glsl_vs.py), its
// interface declarations are rewritten mechanically (uniform / in / out -> plain variables, float literals get an
// `f`), and the body is compiled UNCHANGED inside namespace glsl against the types and built-ins below.
//
// What this header fixes where GLSL leaves the implementation free (all stated in tests/test_vertex_stage_oracle.py):
//   * float is IEEE binary32, no contraction (-ffp-contract=off), cos / sin / pow are libm's cosf / sinf / powf;
//   * mat4 * vec4 is ((c0*x + c1*y) + c2*z) + c3*w;
//   * fract, mod, mix, clamp are the formulas of the GLSL specification (8.3 Common Functions);
//   * textureLod on an UNORM8 image with LINEAR filtering is the weighted sum of OpenGL 3.3 section 3.8.11 with exact
//     float weights (hardware quantises the weights to 8 fractional bits).
// Never compiled into, linked with or loaded by the product (spear_b200/).
#pragma once
#include <math.h>
#include <stdint.h>

namespace glsl {

typedef unsigned int uint;

struct vec2; struct vec3; struct vec4;

// read-only swizzles: a view of N floats that converts to the vector it selects
template <int N, int A, int B> struct Swz2 { float d[N]; inline operator vec2() const; };
template <int N, int A, int B, int C> struct Swz3 { float d[N]; inline operator vec3() const; };

struct vec2 {
	union {
		struct { float x, y; };
		Swz2<2, 0, 1> xy; Swz2<2, 1, 0> yx;
	};
	vec2() : x(0.0f), y(0.0f) { }
	explicit vec2(float s) : x(s), y(s) { }
	vec2(float x_, float y_) : x(x_), y(y_) { }
	vec2(const vec2 &o) : x(o.x), y(o.y) { }
	vec2 &operator=(const vec2 &o) { x = o.x; y = o.y; return *this; }
	vec2 &operator*=(float s) { x *= s; y *= s; return *this; }
	vec2 &operator*=(const vec2 &o) { x *= o.x; y *= o.y; return *this; }
	vec2 &operator+=(const vec2 &o) { x += o.x; y += o.y; return *this; }
	vec2 &operator-=(const vec2 &o) { x -= o.x; y -= o.y; return *this; }
};

struct vec3 {
	union {
		struct { float x, y, z; };
		Swz2<3, 0, 1> xy; Swz2<3, 0, 2> xz; Swz2<3, 1, 2> yz;
		Swz3<3, 0, 1, 2> xyz;
	};
	vec3() : x(0.0f), y(0.0f), z(0.0f) { }
	explicit vec3(float s) : x(s), y(s), z(s) { }
	vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) { }
	vec3(const vec2 &a, float z_) : x(a.x), y(a.y), z(z_) { }
	vec3(const vec3 &o) : x(o.x), y(o.y), z(o.z) { }
	vec3 &operator=(const vec3 &o) { x = o.x; y = o.y; z = o.z; return *this; }
	vec3 &operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
	vec3 &operator*=(const vec3 &o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
	vec3 &operator+=(const vec3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
	vec3 &operator-=(const vec3 &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
};

struct vec4 {
	union {
		struct { float x, y, z, w; };
		Swz2<4, 0, 1> xy; Swz2<4, 2, 3> zw; Swz2<4, 0, 2> xz; Swz2<4, 1, 2> yz;
		Swz3<4, 0, 1, 2> xyz;
	};
	vec4() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) { }
	explicit vec4(float s) : x(s), y(s), z(s), w(s) { }
	vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) { }
	vec4(const vec3 &a, float w_) : x(a.x), y(a.y), z(a.z), w(w_) { }
	vec4(const vec2 &a, const vec2 &b) : x(a.x), y(a.y), z(b.x), w(b.y) { }
	vec4(const vec4 &o) : x(o.x), y(o.y), z(o.z), w(o.w) { }
	vec4 &operator=(const vec4 &o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
	vec4 &operator*=(float s) { x *= s; y *= s; z *= s; w *= s; return *this; }
};

template <int N, int A, int B> inline Swz2<N, A, B>::operator vec2() const { return vec2(d[A], d[B]); }
template <int N, int A, int B, int C> inline Swz3<N, A, B, C>::operator vec3() const { return vec3(d[A], d[B], d[C]); }

// component-wise arithmetic; plain (non-template) functions so that a swizzle converts on either side
#define GLSL_OPS2(OP) \
	inline vec2 operator OP(const vec2 &a, const vec2 &b) { return vec2(a.x OP b.x, a.y OP b.y); } \
	inline vec2 operator OP(const vec2 &a, float b) { return vec2(a.x OP b, a.y OP b); } \
	inline vec2 operator OP(float a, const vec2 &b) { return vec2(a OP b.x, a OP b.y); }
#define GLSL_OPS3(OP) \
	inline vec3 operator OP(const vec3 &a, const vec3 &b) { return vec3(a.x OP b.x, a.y OP b.y, a.z OP b.z); } \
	inline vec3 operator OP(const vec3 &a, float b) { return vec3(a.x OP b, a.y OP b, a.z OP b); } \
	inline vec3 operator OP(float a, const vec3 &b) { return vec3(a OP b.x, a OP b.y, a OP b.z); }
#define GLSL_OPS4(OP) \
	inline vec4 operator OP(const vec4 &a, const vec4 &b) { return vec4(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); } \
	inline vec4 operator OP(const vec4 &a, float b) { return vec4(a.x OP b, a.y OP b, a.z OP b, a.w OP b); } \
	inline vec4 operator OP(float a, const vec4 &b) { return vec4(a OP b.x, a OP b.y, a OP b.z, a OP b.w); }
GLSL_OPS2(+) GLSL_OPS2(-) GLSL_OPS2(*) GLSL_OPS2(/)
GLSL_OPS3(+) GLSL_OPS3(-) GLSL_OPS3(*) GLSL_OPS3(/)
GLSL_OPS4(+) GLSL_OPS4(-) GLSL_OPS4(*) GLSL_OPS4(/)
inline vec2 operator-(const vec2 &a) { return vec2(-a.x, -a.y); }
inline vec3 operator-(const vec3 &a) { return vec3(-a.x, -a.y, -a.z); }
inline vec4 operator-(const vec4 &a) { return vec4(-a.x, -a.y, -a.z, -a.w); }

struct mat4 {
	vec4 c[4]; // columns
	mat4() { }
	mat4(const vec4 &c0, const vec4 &c1, const vec4 &c2, const vec4 &c3) { c[0] = c0; c[1] = c1; c[2] = c2; c[3] = c3; }
};
inline vec4 operator*(const mat4 &m, const vec4 &v) { return ((m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z) + m.c[3] * v.w; }

// GLSL 3.30 specification, 8.1-8.3. The unqualified names hide libm's inside this namespace on purpose.
inline float floor(float x) { return ::floorf(x); }
inline float cos(float x) { return ::cosf(x); }
inline float sin(float x) { return ::sinf(x); }
inline float pow(float x, float y) { return ::powf(x, y); }
inline float min(float a, float b) { return b < a ? b : a; }
inline float max(float a, float b) { return a < b ? b : a; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float fract(float x) { return x - ::floorf(x); }
inline float mod(float x, float y) { return x - y * ::floorf(x / y); }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline vec2 floor(const vec2 &v) { return vec2(floor(v.x), floor(v.y)); }
inline vec3 floor(const vec3 &v) { return vec3(floor(v.x), floor(v.y), floor(v.z)); }
inline vec4 floor(const vec4 &v) { return vec4(floor(v.x), floor(v.y), floor(v.z), floor(v.w)); }
inline vec2 fract(const vec2 &v) { return vec2(fract(v.x), fract(v.y)); }
inline vec3 fract(const vec3 &v) { return vec3(fract(v.x), fract(v.y), fract(v.z)); }
inline vec4 fract(const vec4 &v) { return vec4(fract(v.x), fract(v.y), fract(v.z), fract(v.w)); }
inline vec2 max(const vec2 &v, float s) { return vec2(max(v.x, s), max(v.y, s)); }
inline vec3 max(const vec3 &v, float s) { return vec3(max(v.x, s), max(v.y, s), max(v.z, s)); }
inline vec3 clamp(const vec3 &v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline vec3 mix(const vec3 &x, const vec3 &y, float a) { return x * (1.0f - a) + y * a; }
inline vec4 mix(const vec4 &x, const vec4 &y, float a) { return x * (1.0f - a) + y * a; }

// A two-dimensional UNORM8 image with 1, 2 or 4 channels, GL_LINEAR min / mag filter, no mip chain.
struct sampler2D {
	const uint8_t *texels; int width, height, channels; bool repeat; // repeat: GL_REPEAT, else GL_CLAMP_TO_EDGE
	sampler2D() : texels(0), width(0), height(0), channels(0), repeat(false) { }
	int wrap(int i, int n) const {
		if (repeat) { i %= n; return i < 0 ? i + n : i; }
		return i < 0 ? 0 : i >= n ? n - 1 : i;
	}
	vec4 fetch(int i, int j) const {
		const uint8_t *t = texels + ((size_t)wrap(j, height) * (size_t)width + (size_t)wrap(i, width)) * (size_t)channels;
		float v[4] = { 0.0f, 0.0f, 0.0f, 1.0f }; // missing components read (0, 0, 0, 1)
		for (int k = 0; k < channels; k++) v[k] = (float)t[k] / 255.0f;
		return vec4(v[0], v[1], v[2], v[3]);
	}
};
// OpenGL 3.3 core, 3.8.11 (texture minification, LINEAR): u' = u * w; i0 = floor(u' - 1/2); alpha = frac(u' - 1/2);
// tau = (1-a)(1-b) t00 + a(1-b) t10 + (1-a) b t01 + a b t11
inline vec4 textureLod(const sampler2D &s, const vec2 &uv, float)
{
	const float u = uv.x * (float)s.width - 0.5f, v = uv.y * (float)s.height - 0.5f;
	const float fu = ::floorf(u), fv = ::floorf(v);
	const float a = u - fu, b = v - fv;
	const int i0 = (int)fu, j0 = (int)fv;
	const vec4 t00 = s.fetch(i0, j0), t10 = s.fetch(i0 + 1, j0), t01 = s.fetch(i0, j0 + 1), t11 = s.fetch(i0 + 1, j0 + 1);
	return ((t00 * ((1.0f - a) * (1.0f - b)) + t10 * (a * (1.0f - b))) + t01 * ((1.0f - a) * b)) + t11 * (a * b);
}
inline vec4 texture(const sampler2D &s, const vec2 &uv) { return textureLod(s, uv, 0.0f); } // a vertex shader has no derivatives: lod 0

} // namespace glsl
