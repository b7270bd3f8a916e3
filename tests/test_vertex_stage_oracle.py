"""CPU: oracle/vertex_stage.py (the numpy restatement of Particle.glsl `vs`, src/game/shader/Particle.glsl:73-140)
against the reference's SHIPPED GLSL 3.30 text executed on the CPU (oracle/glsl/: Particle_vs_source_glsl330 of
src/game/shader/Particle.h:239 compiled against a GLSL-semantics header; golden vectors of that run in
tests/golden/vertex_stage.json), and against hand-computed known answers.

Tolerance (floating point, stated): every output is bit-identical to the executed shader except gl_Position.xy
(cos / sin of the rotation) and v_Color.rgb (pow of the sRGB decode), where numpy's float32 kernels and libm differ
by a few ulp: |a - b| <= 4e-6 * max(1, |b|)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import glsl_vs as G
from oracle import vertex_stage as VS

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vertex_stage.json")
TRANSCENDENTAL_COLS = [0, 1, 8, 9, 10]   # gl_Position.xy, v_Color.rgb
EXACT_COLS = [c for c in range(16) if c not in TRANSCENDENTAL_COLS]

F = np.float32
IDENT = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]


def _instance(inv_delta=0.0, aspect=0.5, p2w=IDENT, w2c=IDENT, fog_mad=(0, 0, 0, 0)):
    return np.array(list(p2w) + list(w2c) + list(fog_mad) + [inv_delta, aspect, 0, 0], dtype=F)


def _type(frame_count=(1, 1), frame_rate=0.0, rot=(0, 0, 0, 0), start_frame=0.0, scale=(2.0, 0.0), life=(4.0, 0.0), rel=0.0):
    t = np.zeros(24, dtype=F)
    t[0:2] = frame_count; t[2] = frame_rate; t[4:8] = rot; t[9] = start_frame
    t[12:16] = (F(63) / F(512), F(1) / F(512), F(0.5) / F(512), F(0.5) / F(512))   # u_SplineMad of type 0 (ParticleSystem.cpp:406-413)
    t[16:18] = scale; t[18:20] = life; t[20] = rel
    return t


def _lut(spline=(255, 255, 255, 255), grad=(255, 255, 255, 0)):
    lut = np.zeros((2, 64, 4), dtype=np.uint8)
    lut[0, :] = spline
    lut[1, :] = grad
    return lut.reshape(-1)


def _particle(pos=(1, 2, 3), life=0.5, vel=(0, 0, 0), seed=0.0):
    return np.array([list(pos) + [life] + list(vel) + [seed]], dtype=F)


def test_corners_and_clip_position():
    # rotation 0: axisX = (1, 0), axisY = (0, -1); scale 2 -> corner offsets (+-2, +-2) * (0.5 * aspect, 0.5)
    out = VS.shade(_particle(), _instance(aspect=0.5), _type(), _lut())
    assert out.shape == (4, 16)
    want_xy = [(1 - 0.5, 2 + 1), (1 + 0.5, 2 + 1), (1 - 0.5, 2 - 1), (1 + 0.5, 2 - 1)]  # corners (u, v) = (0,0) (1,0) (0,1) (1,1)
    for c in range(4):
        assert tuple(out[c, 0:2]) == want_xy[c]
        assert tuple(out[c, 2:4]) == (3.0, 1.0)
        assert tuple(out[c, 8:12]) == (1.0, 1.0, 1.0, 1.0)       # white gradient, alpha = spline.y = 1
        assert tuple(out[c, 12:16]) == (0.0, 1.0, 1.0, 0.0)      # frameD, spline.z, erode (spline.w = 1 -> 1), 0


def test_dead_quad_collapses():
    # lifeLeft = life + invDelta / total <= 0 -> gl_Position = 0 (Particle.glsl:113-115)
    out = VS.shade(_particle(life=-0.25), _instance(inv_delta=1.0), _type(life=(4.0, 0.0)), _lut())
    assert np.all(out[:, 0:4] == 0.0)
    out = VS.shade(_particle(life=-0.2), _instance(inv_delta=1.0), _type(life=(4.0, 0.0)), _lut())
    assert np.all(out[:, 3] == 1.0)  # -0.2 + 0.25 > 0: alive


def test_back_extrapolation_and_particle_to_world():
    # world = M * p with a translation (5, 0, 0); position -= velocity * invDelta (:89); velocity ignores translation
    m = list(IDENT); m[12] = 5.0
    out = VS.shade(_particle(pos=(1, 2, 3), vel=(2, 0, -4)), _instance(inv_delta=0.5, p2w=m), _type(scale=(0.0, 0.0)), _lut())
    assert tuple(out[0, 0:4]) == (1 + 5 - 1.0, 2.0, 3 + 2.0, 1.0)


def test_seed_unpack_drives_scale_and_rotation():
    seed = 1 + 64 * 2 + 4096 * 3 + 262144 * 4   # four 6-bit fields (:87): s = (1, 2, 3, 4) / 64
    t = _type(scale=(1.0, 64.0), rot=(0.0, 0.0, 0.0, 0.0))
    out = VS.shade(_particle(seed=float(seed)), _instance(aspect=1.0), t, _lut())
    # scale = 1 + (1/64) * 64 = 2 -> x offset of corner 1 = +2 * 0.5
    assert out[1, 0] - out[0, 0] == 2.0
    # rotation = base + variance * (s.z * 2 - 1) with s.z = 3/64, lifeTime-independent part only
    t = _type(scale=(2.0, 0.0), rot=(0.0, np.pi, 0.0, 0.0))
    out = VS.shade(_particle(seed=float(seed)), _instance(aspect=1.0), t, _lut())
    rot = F(np.pi) * (F(3 / 64) * F(2) - F(1))
    ox, oy = F(-2.0), F(-2.0)
    want = F(1) + (ox * np.cos(rot) + oy * np.sin(rot)) * F(0.5)
    assert abs(out[0, 0] - want) < 1e-6


def test_flipbook_frames():
    # lifeTime 0.5, total 4 s -> absolute 2 s * 8 fps = frame 16 of a 4 x 2 sheet: (0, 0), next (1, 0)
    out = VS.shade(_particle(), _instance(), _type(frame_count=(4, 2), frame_rate=8.0), _lut())
    assert tuple(out[0, 4:8]) == (0.0, 0.0, 0.25, 0.0)
    assert tuple(out[3, 4:8]) == (0.25, 0.5, 0.5, 0.5)
    assert out[0, 12] == 0.0
    # relative frame rate: frame = lifeTime * rate = 0.5 * 5 = 2.5 -> frame 2, blend 0.5
    out = VS.shade(_particle(), _instance(), _type(frame_count=(4, 2), frame_rate=5.0, rel=1.0), _lut())
    assert tuple(out[0, 4:8]) == (0.5, 0.0, 0.75, 0.0) and out[0, 12] == 0.5


def test_lut_lerp_srgb_and_erosion():
    lut = np.zeros((2, 64, 4), dtype=np.uint8)
    lut[0, :, 0] = 255
    lut[0, :, 1] = np.arange(64) * 4          # alpha ramps 0, 4, 8, ...
    lut[0, :, 3] = 51                          # erosion 0.2 -> erode = 5
    lut[1, :, 0] = 128                         # sRGB 128 -> linear 0.2158605
    lut[1, :, 1] = 255
    # lifeTime = 0.5 -> texel coordinate 31.5: alpha = lerp(124, 128, 0.5) / 255
    out = VS.shade(_particle(life=0.5), _instance(), _type(), lut.reshape(-1))
    assert abs(out[0, 11] - 126 / 255) < 1e-7
    assert abs(out[0, 8] - 0.2158605) < 1e-6 and out[0, 9] == 1.0 and out[0, 10] == 0.0
    assert abs(out[0, 14] - 5.0) < 1e-5
    # the ends of the row clamp: lifeTime >= 1 samples texel 63 exactly
    out = VS.shade(_particle(life=-1.0), _instance(), _type(), lut.reshape(-1))
    assert abs(out[0, 11] - 252 / 255) < 1e-7


def test_vis_fog():
    fog = np.zeros((1, 1, 2), dtype=np.uint8)
    fog[0, 0] = (255, 128)       # visible (t = 1), u = 128 / 255 -> colour * u^2
    out = VS.shade(_particle(), _instance(fog_mad=(0.1, 0.1, 0.5, 0.5)), _type(), _lut(), fog=fog)
    assert abs(out[0, 8] - (128 / 255) ** 2) < 1e-6
    fog[0, 0] = (114, 255)       # 114/255 = 0.447 < 0.45 -> t = 0 -> black
    out = VS.shade(_particle(), _instance(fog_mad=(0.1, 0.1, 0.5, 0.5)), _type(), _lut(), fog=fog)
    assert out[0, 8] == 0.0 and out[0, 11] == 1.0   # alpha is not fogged


def _hex(h, dtype, shape=None):
    a = np.frombuffer(bytes.fromhex(h), dtype=dtype)
    return a.reshape(shape) if shape else a


def _golden_cases():
    with open(GOLDEN) as f:
        return json.load(f)


def _case_inputs(c):
    fog = None if c["fog"] is None else _hex(c["fog"]["texels"], np.uint8, (c["fog"]["height"], c["fog"]["width"], 2))
    return (_hex(c["particles"], F, (c["num_particles"], 8)), _hex(c["instance_ubo"], F), _hex(c["type_ubo"], F), _hex(c["lut"], np.uint8), fog,
            _hex(c["vertices"], F, (4 * c["num_particles"], 16)))


def _assert_matches_shader(got, want, what):
    assert np.array_equal(got[:, EXACT_COLS].view(np.uint32), want[:, EXACT_COLS].view(np.uint32)), "%s: an IEEE-only output differs from the executed shader" % what
    ok = np.abs(got - want) <= 4e-6 * np.maximum(1.0, np.abs(want))
    assert ok.all(), "%s: %d floats off, worst %g" % (what, (~ok).sum(), np.abs(got - want).max())


@pytest.mark.parametrize("idx", range(5))
def test_restatement_matches_golden_of_the_executed_shader(idx):
    doc = _golden_cases()
    c = doc["cases"][idx]
    p, inst, typ, lut, fog, want = _case_inputs(c)
    got = VS.shade(p, inst, typ, lut, fog=fog)
    _assert_matches_shader(got, want, c["name"])
    # the cases cover the dead-quad collapse, both erosion branches and the clamped ends of the curve
    if idx == 0:
        all_want = np.concatenate([_case_inputs(k)[5] for k in doc["cases"]])
        assert (np.all(all_want[:, 0:4] == 0.0, axis=1)).any() and (all_want[:, 14] == 1.0).any() and (all_want[:, 14] == 100.0).any()


@pytest.mark.skipif(not G.have(), reason="oracle/_ref/libspear_ref_glslvs.so not built (needs /root/reference)")
def test_executed_shader_reproduces_golden_and_is_the_shipped_text():
    doc = _golden_cases()
    src = G.source()
    assert src.startswith("#version 330") and "a_PositionLife" in src and "u_SplineTexture" in src
    assert hashlib.sha256(src.encode("ascii")).hexdigest() == doc["shader_sha256"] and len(src) == doc["shader_bytes"]
    for c in doc["cases"]:
        p, inst, typ, lut, fog, want = _case_inputs(c)
        atlas = G.atlas_from_luts({c["type_id"]: lut})
        got = G.shade(p, inst, typ, atlas, fog)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), c["name"]


@pytest.mark.skipif(not G.have(), reason="oracle/_ref/libspear_ref_glslvs.so not built (needs /root/reference)")
@pytest.mark.parametrize("type_id", [0, 255, 256, 1300, 2047])
def test_restatement_matches_executed_shader_on_random_input(type_id):
    rng = np.random.default_rng(type_id + 1)
    n = 4000
    lut = rng.integers(0, 256, size=512, dtype=np.uint8)          # worst case: neighbouring texels differ by up to 1.0
    atlas = G.atlas_from_luts({type_id: lut, (type_id + 1) % 2048: rng.integers(0, 256, size=512, dtype=np.uint8),
                               (type_id + 2047) % 2048: rng.integers(0, 256, size=512, dtype=np.uint8)})
    p = np.zeros((n, 8), dtype=F)
    p[:, 0:3] = rng.normal(0, 5, (n, 3)); p[:, 3] = rng.uniform(-0.05, 1.0, n)
    p[:, 4:7] = rng.normal(0, 2, (n, 3)); p[:, 7] = np.floor(rng.uniform(0, 1 << 24, n))
    p[:40, 3] = 1.0; p[40:80, 3] = 0.0
    inst = np.zeros(40, dtype=F)
    inst[0:16] = rng.normal(0, 1, 16); inst[16:32] = rng.normal(0, 1, 16)
    inst[32:36] = (0.02, 0.03, 0.4, 0.3); inst[36] = 0.004; inst[37] = 0.5625
    typ = _type(frame_count=(4, 2), frame_rate=12.0, rot=(0.5, 3.1, 1.5, 0.8), start_frame=1.0, scale=(0.6, 0.4), life=(0.8, F(0.5) * F(1 / 16777216.0)), rel=0.3)
    typ[14] = (F((type_id // 256) * 64) + F(0.5)) / F(512); typ[15] = (F((type_id % 256) * 2) + F(0.5)) / F(512)
    fog = rng.integers(0, 256, size=(24, 32, 2), dtype=np.uint8)
    for fg in (None, fog):
        _assert_matches_shader(VS.shade(p, inst, typ, lut, fog=fg), G.shade(p, inst, typ, atlas, fg), "type %d fog %s" % (type_id, fg is not None))


def _reference_available():
    from oracle import orc
    return orc.have("reference") and G.have()


@pytest.mark.skipif(not _reference_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_a_reference_frame_through_the_reference_shader():
    """Everything from the reference at once, on the CPU: its ParticleSystem.cpp simulates and renders a few frames (headless
    harness), the atlas is the one ITS sg_make_image / sg_bqq_copy_subimages calls filled, the uniform blocks are the ones it
    handed to sg_apply_uniforms, the vertex stream is its own -- and its shipped GLSL runs over that. The restatement, fed
    the type's two LUT rows instead of the atlas, must give the same vertices; the atlas must hold every type's rows in the
    cell oracle/glsl_vs.atlas_from_luts assumes (ParticleSystem.cpp:323-324, :403-404)."""
    import scenarios as S
    from oracle import orc
    from spear_b200 import abi
    from tests_render_args import make_render_args
    camera = (16.0, 20.0, -30.0)
    comps = [
        abi.make_component(
            timeStep=1 / 60, burstAmount=600, spawnTime=0.002, spawnTimeVariance=0.001, lifeTime=0.8, lifeTimeVariance=0.5,
            drag=0.3, gravity=(0, -9.81, 0), size=0.6, sizeVariance=0.4, rotation=30.0, rotationVariance=180.0, spin=90.0, spinVariance=45.0,
            frameRate=12.0, randomStartFrame=1, emitPosition=dict(boxExtent=(1, 0.5, 1)), emitVelocity=dict(sphere=dict(minRadius=1, maxRadius=3)),
            scaleSpline=[(0, 0), (0.3, 1), (1, 0.2)], alphaSpline=[(0, 1), (1, 0)], additiveSpline=[(0, 0.3), (1, 0.9)],
            erosionSpline=[(0, 1), (0.5, 0.4), (1, 0.1)], gradient=dict(points=[(0, (1, 0.5, 0.1)), (0.5, (0.9, 0.1, 0.4)), (1, (0.1, 0.1, 0.1))])),
        abi.make_component(
            timeStep=1 / 50, burstAmount=300, spawnTime=0.01, lifeTime=1.5, lifeTimeVariance=1.0, localSpace=1, size=1.5, relativeFrameRate=1,
            frameRate=8.0, spin=-200.0, emitVelocity=dict(boxExtent=(2, 4, 2)), gradient=dict(defaultColor=(0.2, 0.7, 1.0))),
        S.comp_fountain(),
    ]
    comps[0].frameCount[0], comps[0].frameCount[1] = 4, 2
    comps[1].frameCount[0], comps[1].frameCount[1] = 3, 3
    trs = [abi.make_transform((2, 1, 3), (0, 0.38268343, 0, 0.92387953), 1.5), abi.make_transform((14, 0, 12)), abi.make_transform((20, 2, 18), (0, 0, 0, 1), 0.5)]
    o = orc.OracleSystem("reference", (1, 7, 3, 4))
    try:
        ids = [o.add_effect(c, t) for c, t in zip(comps, trs)]
        ra = make_render_args(camera)
        for k in range(4):      # everything is in view
            for pl in range(4):
                ra.frustum.side[k][pl] = 1.0 if k == 3 else 0.0
                ra.frustum.caps[k][pl] = 1.0 if k == 3 else 0.0
        checked = dead_quads = 0
        for f in range(72):
            o.update(ids, abi.make_frame_args(0.011 + 0.002 * (f % 5), f * 0.017, f, camera=camera))
            if f < 28 or f % 4:
                continue
            recs, _ = o.render(ids, ra, capacity=8)
            assert len(recs) == 3
            atlas = o.read_atlas()
            assert atlas is not None
            luts = {r.typeId: o.type_lut(r.typeId) for r in recs}
            mine = G.atlas_from_luts(luts)
            for a in (atlas, mine):
                a[1::2, :, 3] = 0      # gradient rows: the alpha byte is uninitialised stack in the reference (:326, :366-400), the shader reads .xyz
            assert np.array_equal(atlas, mine), "the reference's atlas does not hold the types' rows where atlas_from_luts puts them"
            for r in recs:
                stream = o.read_stream(r.effectId)
                assert len(stream) == 4 * r.numParticles and r.numParticles > 0
                inst = np.frombuffer(bytes(r.instanceUbo), dtype=F)[:40]
                typ = np.frombuffer(bytes(o.type_ubo(r.typeId)), dtype=F)[:24]
                ref = G.shade(stream[::4], inst, typ, atlas, None)
                _assert_matches_shader(VS.shade(stream[::4], inst, typ, luts[r.typeId], fog=None), ref, "frame %d effect %d" % (f, r.effectId))
                dead_quads += int(np.all(ref[:, 0:4] == 0.0, axis=1).sum()) // 4
                checked += 1
        assert checked == 33 and dead_quads > 0, (checked, dead_quads)   # particles in their last step collapse (Particle.glsl:113-115)
    finally:
        o.close()
