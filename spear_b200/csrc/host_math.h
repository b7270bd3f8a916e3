// host_math.h -- host-side type registration maths of the particle path.
//
// Mirrors what the reference does on the CPU at effect-type registration
// (ParticleSystemImp::initEffectType, ParticleSystem.cpp:298-450;
//  RandomVec3Imp::init, :91-116) and at addEffect (:744-760). These run once
// per type / effect on the host in the reference as well, and use the same
// libm (cosf/sinf/powf) as the reference build on the same machine.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/spear_particles.h"
#include "device_types.h"

namespace spr {

static const float kPi = 3.14159265358979323846f;   // sf::F_PI, src/sf/Base.h:167
static const float kDegToRad = kPi / 180.0f;

// PCG32 stepping on the host (initRng for the per-effect timeStep jitter, :250,:744).
struct HostRng {
	uint64_t state, inc;
	HostRng(uint64_t s = 1, uint64_t i = 1) : state(s), inc(i | 1) { }     // src/sf/Random.h:11
	uint32_t nextU32() {
		uint64_t old = state;
		state = old * 6364136223846793005ULL + inc;
		uint32_t x = (uint32_t)(((old >> 18u) ^ old) >> 27u), rot = (uint32_t)(old >> 59u);
		return (x >> rot) | (x << ((0u - rot) & 31u));
	}
	float nextFloat() { return (float)((double)nextU32() * 2.3283064365386963e-10); }
};

// LCG jump constants for n steps: state' = state*mul + inc*sum, sum = 1 + a + ... + a^(n-1).
inline void lcgJump(uint64_t n, uint64_t &mul, uint64_t &sum)
{
	const uint64_t a = 6364136223846793005ULL;
	uint64_t curMul = a, curSum = 1, accMul = 1, accSum = 0;
	while (n) {
		if (n & 1) { accSum = accSum * curMul + curSum; accMul *= curMul; }
		curSum = (curMul + 1) * curSum;
		curMul *= curMul;
		n >>= 1;
	}
	mul = accMul; sum = accSum;
}

inline float approxCbrtHost(float a) { return 1.5142669123810535f * sqrtf(a) - 0.5142669123810535f * a; } // :62-64

inline void bakeRandomVec3(RandomVec3Dev &d, const SprRandomVec3 &s) // :91-116
{
	memset(&d, 0, sizeof(d));
	d.offset[0] = s.offset.x; d.offset[1] = s.offset.y; d.offset[2] = s.offset.z;
	if (s.boxExtent.x != 0.0f || s.boxExtent.y != 0.0f || s.boxExtent.z != 0.0f) {
		d.flags |= kRvBox;
		d.boxExtent[0] = s.boxExtent.x; d.boxExtent[1] = s.boxExtent.y; d.boxExtent[2] = s.boxExtent.z;
	}
	const SprRandomSphere &sp = s.sphere;
	if (sp.maxRadius > 0.0f) {
		d.flags |= kRvSphere;
		d.thetaBias = sp.minTheta * kDegToRad;
		d.thetaScale = (sp.maxTheta - sp.minTheta) * kDegToRad;
		d.cosPhiBias = cosf(sp.maxPhi * kDegToRad);
		d.cosPhiScale = cosf(sp.minPhi * kDegToRad) - d.cosPhiBias;
		d.radiusCubeScale = approxCbrtHost(1.0f - sp.minRadius / sp.maxRadius);
		d.radiusScale = sp.maxRadius;
		d.sphereScale[0] = sp.scale.x; d.sphereScale[1] = sp.scale.y; d.sphereScale[2] = sp.scale.z;
	}
	if (s.rotation.x != 0.0f || s.rotation.y != 0.0f || s.rotation.z != 0.0f) {
		d.flags |= kRvRotation;
		// sf::eulerAnglesToQuat, XYZ order (src/sf/Quaternion.cpp:18-33)
		float hx = s.rotation.x * kDegToRad, hy = s.rotation.y * kDegToRad, hz = s.rotation.z * kDegToRad;
		hx *= 0.5f; hy *= 0.5f; hz *= 0.5f;
		float cx = cosf(hx), sx = sinf(hx), cy = cosf(hy), sy = sinf(hy), cz = cosf(hz), sz = sinf(hz);
		d.rotation[0] = -cx*sy*sz + cy*cz*sx;
		d.rotation[1] = cx*cz*sy + cy*sx*sz;
		d.rotation[2] = cx*cy*sz - cz*sx*sy;
		d.rotation[3] = cx*cy*cz + sx*sy*sz;
	}
}

inline uint32_t drawsOfSampler(const RandomVec3Dev &d) { return ((d.flags & kRvBox) ? 3u : 0u) + ((d.flags & kRvSphere) ? 3u : 0u); }

// Uniform cubic B-spline segment basis at parameter t (src/client/BSpline.cpp:7-35);
// each weight is the exact float expression the reference multiplies the control point by.
struct BSplineBasis {
	float w[4];
	explicit BSplineBasis(float t) {
		float nt = 1.0f - t, t2 = t*t, t3 = t2*t, nt2 = nt*nt, nt3 = nt2*nt;
		w[0] = nt3 / 6.0f;
		w[1] = (3.0f*t3 - 6.0f*t2 + 4.0f) / 6.0f;
		w[2] = (-3.0f*t3 + 3.0f*t2 + 3.0f*t + 1.0f) / 6.0f;
		w[3] = t3 / 6.0f;
	}
	float blend(float a, float b, float c, float d) const { float p = a * w[0]; p += b * w[1]; p += c * w[2]; p += d * w[3]; return p; }
};

// cl::discretizeBSplineY (src/client/BSpline.cpp:37-88): y at `count` even x positions,
// inverting x(t) per segment with a 512-step lower-bound bisection.
inline void sampleBSplineY(float *y, uint32_t count, const SprVec2 *pts, uint32_t numPts, float xMin, float xMax)
{
	if (count == 0) return;
	if (count == 1) { y[0] = pts[0].y; return; }
	const int n = (int)numPts;
	auto ctrl = [&](int i) -> const SprVec2& { return pts[i < 0 ? 0 : (i >= n ? n - 1 : i)]; };
	const float xStep = (xMax - xMin) / (float)(count - 1);
	const float tStep = 1.0f / 511.0f;
	float x = xMin;
	uint32_t out = 0;
	for (int seg = -1; seg < n && out < count; seg++) {
		const SprVec2 &a = ctrl(seg - 1), &b = ctrl(seg), &c = ctrl(seg + 1), &d = ctrl(seg + 2);
		const float segMaxX = (b.x + d.x + 4.0f * c.x) * (1.0f / 6.0f);
		for (; x <= segMaxX; x += xStep) {
			uint32_t lo = 0, len = 512;
			while (len > 0) {
				uint32_t half = len >> 1, mid = lo + half;
				if (BSplineBasis((float)mid * tStep).blend(a.x, b.x, c.x, d.x) < x) { lo = mid; len -= half + 1; }
				else len = half;
			}
			y[out] = BSplineBasis((float)lo * tStep).blend(a.y, b.y, c.y, d.y);
			if (++out == count) return;
		}
	}
	for (; out < count; out++) y[out] = out > 0 ? y[out - 1] : 0.0f;
}

inline float clamp01(float v) { float m = v > 0.0f ? v : 0.0f; return m < 1.0f ? m : 1.0f; } // sf::clamp, src/sf/Base.h:118-120

inline uint8_t linearToSrgb8(float x) // sp::linearToSrgbUnorm, src/sp/Srgb.cpp:10-22,:48-53
{
	x = clamp01(x);
	float s = x <= 0.0031308f ? 12.92f * x : 1.055f * powf(x, (1.0f / 2.4f)) - 0.055f;
	return (uint8_t)(s * 255.9f);
}

inline void bakeSplineRow(uint8_t *lut, int channel, const SprBSpline2 &sp, bool complement) // :329-364
{
	float y[SPR_SPLINE_SAMPLE_RATE];
	if (sp.numPoints >= 1) sampleBSplineY(y, SPR_SPLINE_SAMPLE_RATE, sp.points, sp.numPoints, 0.0f, 1.0f);
	for (int i = 0; i < SPR_SPLINE_SAMPLE_RATE; i++) {
		float v = sp.numPoints >= 1 ? (complement ? 1.0f - y[i] : y[i]) : 1.0f;
		int q = (int)(v * 255.0f);
		lut[i * 4 + channel] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
	}
}

inline void bakeGradientRow(uint8_t *row, const SprGradient &grad) // :366-401
{
	SprGradientPoint fallback = { 1.0f, grad.defaultColor };
	const SprGradientPoint *pts = grad.numPoints ? grad.points : &fallback;
	const uint32_t n = grad.numPoints ? grad.numPoints : 1;
	const float step = 1.0f / (float)(SPR_SPLINE_SAMPLE_RATE - 1);
	uint32_t next = 0;
	for (uint32_t i = 0; i < SPR_SPLINE_SAMPLE_RATE; i++) {
		float t = (float)i * step;
		while (next < n && t >= pts[next].t) next++;
		SprVec3 col;
		if (next == 0) col = pts[0].color;
		else if (next == n) col = pts[n - 1].color;
		else {
			const SprGradientPoint &p0 = pts[next - 1], &p1 = pts[next];
			float u = (t - p0.t) / (p1.t - p0.t), iu = 1.0f - u; // sf::lerp: a*(1-t) + b*t
			col.x = p0.color.x * iu + p1.color.x * u;
			col.y = p0.color.y * iu + p1.color.y * u;
			col.z = p0.color.z * iu + p1.color.z * u;
		}
		row[i * 4 + 0] = linearToSrgb8(col.x);
		row[i * 4 + 1] = linearToSrgb8(col.y);
		row[i * 4 + 2] = linearToSrgb8(col.z);
		row[i * 4 + 3] = 0; // never written by the reference (:398-399)
	}
}

// Spline atlas addressing (:253-258, :323-324, :403-413)
inline void bakeTypeUbo(SprVertexTypeUbo &u, uint32_t typeId, const SprParticleSystemComponent &c) // :312-321,:410-419
{
	memset(&u, 0, sizeof(u));
	u.u_FrameCount.x = (float)c.frameCount[0];
	u.u_FrameCount.y = (float)c.frameCount[1];
	u.u_FrameRate = c.frameRate;
	u.u_RotationControl.x = c.rotation * kDegToRad;
	u.u_RotationControl.y = c.rotationVariance * kDegToRad;
	u.u_RotationControl.z = c.spin * kDegToRad;
	u.u_RotationControl.w = c.spinVariance * kDegToRad;
	if (c.randomStartFrame) u.u_StartFrame = 1.0f;
	const uint32_t atlasHeight = kSplineAtlasTypesY, atlasWidth = kSplineAtlasTypesX, rowsPerType = kSplineRowsPerType;
	static_assert(SPR_SPLINE_SAMPLE_RATE == kSplineSamples, "one sample rate");
	const float resX = (float)(SPR_SPLINE_SAMPLE_RATE * atlasWidth), resY = (float)(atlasHeight * rowsPerType);
	uint32_t px = (typeId / atlasHeight) * SPR_SPLINE_SAMPLE_RATE, py = (typeId % atlasHeight) * rowsPerType;
	u.u_SplineMad.x = (float)(SPR_SPLINE_SAMPLE_RATE - 1) / resX;
	u.u_SplineMad.y = 1.0f / resY;
	u.u_SplineMad.z = ((float)px + 0.5f) / resX;
	u.u_SplineMad.w = ((float)py + 0.5f) / resY;
	u.u_ScaleBaseVariance.x = c.size;
	u.u_ScaleBaseVariance.y = c.sizeVariance;
	u.u_LifeTimeBaseVariance.x = c.lifeTime;
	u.u_LifeTimeBaseVariance.y = c.lifeTimeVariance * (1.0f / 16777216.0f);
}

// sf::mat::world (src/sf/Matrix.cpp:899-920) == cl::Transform::asMatrix (src/client/Transform.h:15-17)
inline SprMat34 transformToMatrix(const SprTransform &t)
{
	const SprQuat &q = t.rotation;
	float s2 = 2.0f * t.scale;
	float xx = q.x*q.x, xy = q.x*q.y, xz = q.x*q.z, xw = q.x*q.w;
	float yy = q.y*q.y, yz = q.y*q.z, yw = q.y*q.w;
	float zz = q.z*q.z, zw = q.z*q.w;
	SprMat34 m;
	m.v[0] = s2 * (- yy - zz + 0.5f); m.v[1] = s2 * (+ xy + zw); m.v[2] = s2 * (- yw + xz);
	m.v[3] = s2 * (- zw + xy); m.v[4] = s2 * (- xx - zz + 0.5f); m.v[5] = s2 * (+ xw + yz);
	m.v[6] = s2 * (+ xz + yw); m.v[7] = s2 * (- xw + yz); m.v[8] = s2 * (- xx - yy + 0.5f);
	m.v[9] = t.position.x; m.v[10] = t.position.y; m.v[11] = t.position.z;
	return m;
}

inline SprMat34 identity34() { SprMat34 m; memset(&m, 0, sizeof(m)); m.v[0] = m.v[4] = m.v[8] = 1.0f; return m; }

inline void writeColMajor44(const SprMat34 &m, float *dst) // src/sf/Matrix.h:131-136
{
	for (int c = 0; c < 4; c++) {
		dst[c * 4 + 0] = m.v[c * 3 + 0]; dst[c * 4 + 1] = m.v[c * 3 + 1]; dst[c * 4 + 2] = m.v[c * 3 + 2];
		dst[c * 4 + 3] = c == 3 ? 1.0f : 0.0f;
	}
}

// sf::transformBounds (src/sf/Geometry.cpp:13-19, src/sf/Matrix.cpp:579-586,:625-632)
inline SprBounds3 transformBounds(const SprMat34 &t, const SprBounds3 &b)
{
	const float *m = t.v;
	SprBounds3 r;
	r.origin.x = m[0]*b.origin.x + m[3]*b.origin.y + m[6]*b.origin.z + m[9];
	r.origin.y = m[1]*b.origin.x + m[4]*b.origin.y + m[7]*b.origin.z + m[10];
	r.origin.z = m[2]*b.origin.x + m[5]*b.origin.y + m[8]*b.origin.z + m[11];
	r.extent.x = fabsf(m[0])*b.extent.x + fabsf(m[3])*b.extent.y + fabsf(m[6])*b.extent.z;
	r.extent.y = fabsf(m[1])*b.extent.x + fabsf(m[4])*b.extent.y + fabsf(m[7])*b.extent.z;
	r.extent.z = fabsf(m[2])*b.extent.x + fabsf(m[5])*b.extent.y + fabsf(m[8])*b.extent.z;
	return r;
}

// sf::Frustum::intersects(Bounds3) (src/sf/Frustum.h:60-75)
inline bool frustumIntersects(const SprFrustum &f, const SprBounds3 &b)
{
	const float o[3] = { b.origin.x, b.origin.y, b.origin.z }, e[3] = { b.extent.x, b.extent.y, b.extent.z };
	for (int lane = 0; lane < 4; lane++) {
		float so = f.side[3][lane], co = f.caps[3][lane];
		for (int k = 0; k < 3; k++) {
			float s = f.side[k][lane], c = f.caps[k][lane];
			so += s * o[k]; co += c * o[k];
			so += fabsf(s) * e[k]; co += fabsf(c) * e[k];
		}
		float m = so < co ? so : co;
		if (!(m > 0.0f)) return false;
	}
	return true;
}

} // namespace spr
