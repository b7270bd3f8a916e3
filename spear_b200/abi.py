"""ctypes mirror of include/spear_particles.h (POD structs only).

These definitions are shared by the product binding (spear_b200.particles) and
by the test-side oracle loader (oracle/orc.py), so both sides receive
byte-identical descriptors.  Field order and types follow the header, which in
turn follows the reference structs it cites (sv::ParticleSystemComponent,
src/server/ServerState.h:144-239 etc.).
"""
import ctypes as C

SPLINE_LUT_BYTES = 512
MAX_GRAVITY_POINTS = 16


class Vec2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Vec4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class Quat(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class Mat34(C.Structure):
    _fields_ = [("v", C.c_float * 12)]


class Mat44(C.Structure):
    _fields_ = [("v", C.c_float * 16)]


class Transform(C.Structure):
    _fields_ = [("position", Vec3), ("rotation", Quat), ("scale", C.c_float)]


class Bounds3(C.Structure):
    _fields_ = [("origin", Vec3), ("extent", Vec3)]


class Frustum(C.Structure):
    _fields_ = [("side", (C.c_float * 4) * 4), ("caps", (C.c_float * 4) * 4)]


class BSpline2(C.Structure):
    _fields_ = [("points", C.POINTER(Vec2)), ("numPoints", C.c_uint32)]


class GradientPoint(C.Structure):
    _fields_ = [("t", C.c_float), ("color", Vec3)]


class Gradient(C.Structure):
    _fields_ = [("defaultColor", Vec3), ("points", C.POINTER(GradientPoint)), ("numPoints", C.c_uint32)]


class RandomSphere(C.Structure):
    _fields_ = [("minTheta", C.c_float), ("maxTheta", C.c_float), ("minPhi", C.c_float), ("maxPhi", C.c_float),
                ("minRadius", C.c_float), ("maxRadius", C.c_float), ("scale", Vec3)]


class RandomVec3(C.Structure):
    _fields_ = [("offset", Vec3), ("boxExtent", Vec3), ("sphere", RandomSphere), ("rotation", Vec3)]


class GravityPoint(C.Structure):
    _fields_ = [("position", Vec3), ("radius", C.c_float), ("strength", C.c_float)]


class ParticleSystemComponent(C.Structure):
    _fields_ = [
        ("texture", C.c_uint64),
        ("frameCount", C.c_int32 * 2),
        ("timeStep", C.c_float),
        ("prewarmTime", C.c_float),
        ("updateRadius", C.c_float),
        ("cullPadding", C.c_float),
        ("updateOutOfCamera", C.c_uint8),
        ("renderOrder", C.c_double),
        ("spawnTime", C.c_float),
        ("spawnTimeVariance", C.c_float),
        ("burstAmount", C.c_uint32),
        ("burstAmountVariance", C.c_uint32),
        ("emitterOnTime", C.c_float),
        ("localSpace", C.c_uint8),
        ("instantDelete", C.c_uint8),
        ("emitPosition", RandomVec3),
        ("emitVelocity", RandomVec3),
        ("emitVelocityAttractorOffset", Vec3),
        ("emitVelocityAttractorStrength", C.c_float),
        ("gravityPoints", C.POINTER(GravityPoint)),
        ("numGravityPoints", C.c_uint32),
        ("drag", C.c_float),
        ("gravity", Vec3),
        ("size", C.c_float),
        ("sizeVariance", C.c_float),
        ("lifeTime", C.c_float),
        ("lifeTimeVariance", C.c_float),
        ("scaleSpline", BSpline2),
        ("alphaSpline", BSpline2),
        ("additiveSpline", BSpline2),
        ("erosionSpline", BSpline2),
        ("gradient", Gradient),
        ("frameRate", C.c_float),
        ("relativeFrameRate", C.c_uint8),
        ("randomStartFrame", C.c_uint8),
        ("rotation", C.c_float),
        ("rotationVariance", C.c_float),
        ("spin", C.c_float),
        ("spinVariance", C.c_float),
    ]


class GpuParticle(C.Structure):
    _fields_ = [("position", Vec3), ("life", C.c_float), ("velocity", Vec3), ("seed", C.c_float)]


class VertexTypeUbo(C.Structure):
    _fields_ = [("u_FrameCount", Vec2), ("u_FrameRate", C.c_float), ("_pad_12", C.c_uint8 * 4),
                ("u_RotationControl", Vec4), ("u_Additive", C.c_float), ("u_StartFrame", C.c_float),
                ("_pad_40", C.c_uint8 * 8), ("u_SplineMad", Vec4), ("u_ScaleBaseVariance", Vec2),
                ("u_LifeTimeBaseVariance", Vec2), ("u_RelativeFrameRate", C.c_float), ("_pad_84", C.c_uint8 * 12)]


class VertexInstanceUbo(C.Structure):
    _fields_ = [("u_ParticleToWorld", C.c_float * 16), ("u_WorldToClip", C.c_float * 16), ("u_VisFogMad", Vec4),
                ("u_InvDelta", C.c_float), ("u_Aspect", C.c_float), ("_pad_152", C.c_uint8 * 8)]


class SystemsDesc(C.Structure):
    _fields_ = [("persistCamera", C.c_float * 2), ("persistZoom", C.c_float), ("seed", C.c_uint64 * 4)]


class TransformUpdate(C.Structure):
    _fields_ = [("previousTransform", Transform), ("transform", Transform), ("entityToWorld", Mat34)]


class RenderArgs(C.Structure):
    _fields_ = [("cameraPosition", Vec3), ("viewToClip", Mat44), ("worldToClip", Mat44), ("frustum", Frustum),
                ("visFogWorldMad", Vec4)]


class FrameArgs(C.Structure):
    _fields_ = [("frameIndex", C.c_uint64), ("gameTime", C.c_double), ("dt", C.c_float), ("cameraPosition", Vec3)]


class DrawRecord(C.Structure):
    _fields_ = [("effectId", C.c_uint32), ("typeId", C.c_uint32), ("numParticles", C.c_uint32),
                ("typeUboChanged", C.c_uint32), ("vertexByteOffset", C.c_uint64), ("uploadedThisFrame", C.c_uint32),
                ("uploadRingSlot", C.c_uint32), ("instanceUbo", VertexInstanceUbo)]


class ShadedVertex(C.Structure):
    """SprShadedVertex: the outputs of Particle.glsl `vs` for one vertex (64 bytes)."""
    _fields_ = [("position", C.c_float * 4), ("uv0", C.c_float * 2), ("uv1", C.c_float * 2),
                ("color", C.c_float * 4), ("params", C.c_float * 4)]


class FogImage(C.Structure):
    _fields_ = [("deviceTexels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class Sprite(C.Structure):
    """SprSprite: the sp::Sprite fields cl::BillboardSystem reads."""
    _fields_ = [("atlas", C.c_uint64), ("minVert", Vec2), ("maxVert", Vec2), ("vertUvScale", Vec2), ("vertUvBias", Vec2)]


class Billboard(C.Structure):
    """SprBillboard = cl::Billboard (src/client/BillboardSystem.h:9-18)."""
    _fields_ = [("sprite", Sprite), ("transform", C.c_float * 12), ("color", Vec4), ("depth", C.c_float),
                ("anchor", Vec2), ("cropMin", Vec2), ("cropMax", Vec2)]


class BillboardPacked(C.Structure):
    """SprBillboardPacked: the billboard as the device reads it (120 bytes)."""
    _fields_ = [("atlas", C.c_uint64), ("minVert", C.c_float * 2), ("maxVert", C.c_float * 2), ("vertUvScale", C.c_float * 2),
                ("vertUvBias", C.c_float * 2), ("transform", C.c_float * 12), ("anchor", C.c_float * 2), ("color", C.c_uint32),
                ("cropMin", C.c_float * 2), ("cropMax", C.c_float * 2), ("depth", C.c_float)]


class BillboardVertex(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("texCoord", C.c_float * 2), ("color", C.c_uint32)]


class BillboardDraw(C.Structure):
    _fields_ = [("atlas", C.c_uint64), ("count", C.c_uint32), ("firstQuad", C.c_uint32)]


class BillboardOutput(C.Structure):
    _fields_ = [("deviceVertices", C.c_void_p), ("numQuads", C.c_uint32), ("numDraws", C.c_uint32),
                ("draws", C.POINTER(BillboardDraw)), ("worldToClip", C.c_float * 16)]


BILLBOARD_VERTEX_DTYPE = [("position", "<f4", 3), ("texCoord", "<f4", 2), ("color", "<u4")]


def make_billboard(atlas=1, min_vert=(0.0, 0.0), max_vert=(1.0, 1.0), uv_scale=(0.25, 0.25), uv_bias=(0.5, 0.25),
                   transform=(1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0), color=(1, 1, 1, 1), depth=0.0,
                   anchor=(0.5, 0.5), crop_min=(0.0, 0.0), crop_max=(1.0, 1.0)):
    b = Billboard()
    b.sprite = Sprite(int(atlas), Vec2(*map(float, min_vert)), Vec2(*map(float, max_vert)), Vec2(*map(float, uv_scale)), Vec2(*map(float, uv_bias)))
    for i, v in enumerate(transform):
        b.transform[i] = float(v)
    b.color = Vec4(*map(float, color))
    b.depth = float(depth)
    b.anchor = Vec2(*map(float, anchor))
    b.cropMin = Vec2(*map(float, crop_min))
    b.cropMax = Vec2(*map(float, crop_max))
    return b


class RenderOutput(C.Structure):
    _fields_ = [("deviceVertices", C.c_void_p), ("records", C.POINTER(DrawRecord)), ("numRecords", C.c_uint32),
                ("numSortedEffects", C.c_uint32)]


class EffectInfo(C.Structure):
    _fields_ = [("typeId", C.c_uint32), ("slotCount", C.c_uint32), ("aliveCount", C.c_uint32),
                ("freeCount", C.c_uint32), ("timeStep", C.c_float), ("timeDelta", C.c_float),
                ("spawnTimer", C.c_float), ("emitOnTimer", C.c_float), ("stopEmit", C.c_uint8),
                ("firstEmit", C.c_uint8), ("instantDelete", C.c_uint8), ("_pad", C.c_uint8),
                ("gpuBounds", Bounds3), ("lastUpdateTime", C.c_double), ("simSteps", C.c_uint64)]


class ContextDesc(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_uint32), ("maxParticlesPerFrame", C.c_uint32),
                ("maxParticleCacheFrames", C.c_uint32), ("initialSlotCapacity", C.c_uint64),
                ("stream", C.c_void_p)]


class EffectDeviceView(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("slots", C.c_void_p), ("aliveCount", C.c_void_p),
                ("vertexByteOffset", C.c_uint64), ("slotCapacity", C.c_uint32)]


CTX_SORT_PARTICLES = 1


# ---------------------------------------------------------------------------
# Descriptor construction helpers
# ---------------------------------------------------------------------------

def _v3(v):
    return Vec3(float(v[0]), float(v[1]), float(v[2]))


def default_component():
    """sv::ParticleSystemComponent default member initialisers (ServerState.h:187-239)."""
    c = ParticleSystemComponent()
    c.frameCount[0] = 1
    c.frameCount[1] = 1
    c.timeStep = 0.1
    c.updateRadius = 10.0
    c.cullPadding = 1.5
    c.spawnTime = 0.2
    c.emitterOnTime = -1.0
    for rv in (c.emitPosition, c.emitVelocity):
        rv.sphere.maxTheta = 360.0
        rv.sphere.maxPhi = 180.0
        rv.sphere.scale = Vec3(1.0, 1.0, 1.0)
    c.emitVelocityAttractorStrength = 0.0
    c.size = 0.5
    c.lifeTime = 1.0
    c.gradient.defaultColor = Vec3(1.0, 1.0, 1.0)
    return c


_SPLINES = ("scaleSpline", "alphaSpline", "additiveSpline", "erosionSpline")


def make_component(**kw):
    """Build a ParticleSystemComponent from keyword overrides.

    Nested random-vector fields use dicts:
        emitPosition=dict(offset=(..), boxExtent=(..), rotation=(..),
                          sphere=dict(minRadius=.., maxRadius=.., scale=(..), ...))
    splines are lists of (x, y); gradient is dict(defaultColor=.., points=[(t,(r,g,b)),..]);
    gravityPoints is a list of dict(position=.., radius=.., strength=..).
    The returned object keeps the backing arrays alive in `_keep`.
    """
    c = default_component()
    keep = []
    for k, v in kw.items():
        if k in ("emitPosition", "emitVelocity"):
            rv = getattr(c, k)
            for kk, vv in v.items():
                if kk == "sphere":
                    for sk, sv_ in vv.items():
                        if sk == "scale":
                            rv.sphere.scale = _v3(sv_)
                        else:
                            setattr(rv.sphere, sk, float(sv_))
                else:
                    setattr(rv, kk, _v3(vv))
        elif k in _SPLINES:
            arr = (Vec2 * max(len(v), 1))(*[Vec2(float(p[0]), float(p[1])) for p in v])
            keep.append(arr)
            sp = getattr(c, k)
            sp.points = C.cast(arr, C.POINTER(Vec2))
            sp.numPoints = len(v)
        elif k == "gradient":
            if "defaultColor" in v:
                c.gradient.defaultColor = _v3(v["defaultColor"])
            pts = v.get("points", [])
            arr = (GradientPoint * max(len(pts), 1))(*[GradientPoint(float(t), _v3(col)) for t, col in pts])
            keep.append(arr)
            c.gradient.points = C.cast(arr, C.POINTER(GradientPoint))
            c.gradient.numPoints = len(pts)
        elif k == "gravityPoints":
            arr = (GravityPoint * max(len(v), 1))(*[
                GravityPoint(_v3(p["position"]), float(p.get("radius", 0.5)), float(p.get("strength", 1.0)))
                for p in v])
            keep.append(arr)
            c.gravityPoints = C.cast(arr, C.POINTER(GravityPoint))
            c.numGravityPoints = len(v)
        elif k in ("gravity", "emitVelocityAttractorOffset"):
            setattr(c, k, _v3(v))
        elif k == "frameCount":
            c.frameCount[0] = int(v[0])
            c.frameCount[1] = int(v[1])
        else:
            setattr(c, k, v)
    c._keep = keep
    return c


def make_transform(position=(0.0, 0.0, 0.0), rotation=(0.0, 0.0, 0.0, 1.0), scale=1.0):
    return Transform(_v3(position), Quat(*[float(x) for x in rotation]), float(scale))


def make_systems_desc(seed=(1, 2, 3, 4)):
    d = SystemsDesc()
    for i in range(4):
        d.seed[i] = int(seed[i])
    return d


def make_frame_args(dt, game_time=0.0, frame_index=0, camera=(0.0, 0.0, 0.0)):
    return FrameArgs(int(frame_index), float(game_time), float(dt), _v3(camera))
