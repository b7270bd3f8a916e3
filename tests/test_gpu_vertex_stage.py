"""GPU: sprShadeVertices (spear_b200/csrc/vertex_stage.cu) on the draw records of a rendered frame
against (1) the reference's SHIPPED GLSL 3.30 vertex shader executed on the CPU (oracle/_ref/libspear_ref_glslvs.so:
Particle_vs_source_glsl330 of src/game/shader/Particle.h:239 through oracle/glsl/glsl_shim.h, sampling the 512 x 512
spline atlas the reference builds) and (2) the numpy restatement of Particle.glsl `vs` (oracle/vertex_stage.py).

Tolerance (stated, floating point): everything the shader computes with IEEE add / mul / div only -- gl_Position.zw,
flip-book uvs, frame blend, alpha, additive, erosion, the dead-quad collapse -- must be bit-identical (the library is
built with -fmad=false); gl_Position.xy and the gradient colour go through cos / sin / pow, whose CUDA, libm and numpy
implementations differ by a few ulp: |a - b| <= 2e-5 * max(1, |b|)."""
import ctypes as C

import numpy as np
import pytest

import scenarios as S
from oracle import glsl_vs as G
from oracle import vertex_stage as VS
from spear_b200 import abi
from tests_render_args import make_render_args

pytestmark = pytest.mark.gpu

CAMERA = (16.0, 20.0, -30.0)


def _particles():
    from spear_b200 import particles
    return particles


def _comps():
    a = abi.make_component(
        timeStep=1 / 60, burstAmount=3000, spawnTime=0.002, spawnTimeVariance=0.001, lifeTime=0.8, lifeTimeVariance=0.5,
        drag=0.3, gravity=(0, -9.81, 0), size=0.6, sizeVariance=0.4, rotation=30.0, rotationVariance=180.0, spin=90.0, spinVariance=45.0,
        frameRate=12.0, randomStartFrame=1, emitPosition=dict(boxExtent=(1, 0.5, 1)), emitVelocity=dict(sphere=dict(minRadius=1, maxRadius=3)),
        scaleSpline=[(0, 0), (0.3, 1), (1, 0.2)], alphaSpline=[(0, 1), (1, 0)], additiveSpline=[(0, 0.3), (1, 0.9)],
        erosionSpline=[(0, 1), (0.5, 0.4), (1, 0.1)], gradient=dict(points=[(0, (1, 0.5, 0.1)), (0.5, (0.9, 0.1, 0.4)), (1, (0.1, 0.1, 0.1))]))
    a.frameCount[0], a.frameCount[1] = 4, 2
    b = abi.make_component(
        timeStep=1 / 50, burstAmount=700, spawnTime=0.01, lifeTime=1.5, lifeTimeVariance=1.0, localSpace=1, size=1.5, relativeFrameRate=1,
        frameRate=8.0, spin=-200.0, emitVelocity=dict(boxExtent=(2, 4, 2)), gradient=dict(defaultColor=(0.2, 0.7, 1.0)))
    b.frameCount[0], b.frameCount[1] = 3, 3
    c = S.comp_fountain()
    return [a, b, c]


def _close(got, want):
    return np.abs(got - want) <= 2e-5 * np.maximum(1.0, np.abs(want))


@pytest.mark.parametrize("sort", [False, True], ids=["stream", "sorted"])
@pytest.mark.parametrize("with_fog", [False, True], ids=["nofog", "fog"])
def test_shade_vertices_matches_glsl_restatement(sort, with_fog):
    import torch
    P = _particles()
    ctx = P.ParticleContext(sort_particles=sort, initial_slot_capacity=1 << 14, max_particles_per_frame=1 << 20)
    try:
        s = ctx.create_system((1, 7, 3, 4))
        comps = _comps()
        tr = [abi.make_transform((2, 1, 3), (0, 0.38268343, 0, 0.92387953), 1.5), abi.make_transform((14, 0, 12)), abi.make_transform((20, 2, 18), (0, 0, 0, 1), 0.5)]
        ids = [s.add_effect(c, t) for c, t in zip(comps, tr)]
        for f in range(40):
            s.update(ids, abi.make_frame_args(0.017, f * 0.017, f, camera=CAMERA))
        ra = make_render_args(CAMERA)
        ra.visFogWorldMad = abi.Vec4(0.02, 0.03, 0.4, 0.3)
        for k in range(4):      # everything is in view
            for pl in range(4):
                ra.frustum.side[k][pl] = 1.0 if k == 3 else 0.0
                ra.frustum.caps[k][pl] = 1.0 if k == 3 else 0.0
        fog_np, fog = None, None
        if with_fog:
            rng = np.random.default_rng(5)
            fog_np = rng.integers(0, 256, size=(24, 32, 2), dtype=np.uint8)
            fog_t = torch.from_numpy(fog_np).cuda()
            fog = abi.FogImage(fog_t.data_ptr(), 32, 24)
        dead_quads = 0
        shader_checked = 0
        atlas = None
        # several frames with varying dt: u_InvDelta = timeStep - timeDelta changes every frame, and with it which
        # of the particles in their last step collapse (lifeLeft <= 0, Particle.glsl:113-115)
        for f in range(40, 52):
            s.update(ids, abi.make_frame_args(0.011 + 0.002 * (f % 5), f * 0.017, f, camera=CAMERA))
            recs, _ = s.render(ids, ra)
            assert len(recs) == 3
            total = sum(r.numParticles for r in recs)
            out = torch.full((4 * total + 8, 16), -777.0, dtype=torch.float32, device="cuda")
            n = s.shade_vertices(recs, out.data_ptr(), 4 * total, fog)
            ctx.synchronize(); torch.cuda.synchronize()
            assert n == 4 * total
            got = out.cpu().numpy()
            assert np.all(got[4 * total:] == -777.0), "wrote past the vertices it reported"
            v0 = 0
            for r in recs:
                stream = s.read_stream(r.effectId)
                assert len(stream) == 4 * r.numParticles
                inst = np.frombuffer(bytes(r.instanceUbo), dtype=np.float32)[:40]
                typ = np.frombuffer(bytes(s.type_ubo(r.typeId)), dtype=np.float32)[:24]
                want = VS.shade(stream[::4], inst, typ, s.type_lut(r.typeId), fog=fog_np)
                g = got[v0:v0 + 4 * r.numParticles]
                exact_cols = [2, 3, 4, 5, 6, 7, 11, 12, 13, 14, 15]   # gl_Position.zw, uv0, uv1, alpha, frameD, additive, erode, 0
                assert np.array_equal(g[:, exact_cols].view(np.uint32), want[:, exact_cols].view(np.uint32)), "frame %d effect %d: IEEE-only outputs differ" % (f, r.effectId)
                ok = _close(g, want)
                assert ok.all(), "frame %d effect %d: %d of %d floats off, worst %g" % (f, r.effectId, (~ok).sum(), ok.size, np.abs(g - want).max())
                if G.have():
                    # the reference's own shader text, sampling the atlas with every live type in its cell
                    if atlas is None:
                        atlas = G.atlas_from_luts({rr.typeId: s.type_lut(rr.typeId) for rr in recs})
                    ref = G.shade(stream[::4], inst, typ, atlas, fog_np)
                    assert np.array_equal(g[:, exact_cols].view(np.uint32), ref[:, exact_cols].view(np.uint32)), "frame %d effect %d: IEEE-only outputs differ from the executed GLSL" % (f, r.effectId)
                    ok = _close(g, ref)
                    assert ok.all(), "frame %d effect %d against the executed GLSL: %d floats off, worst %g" % (f, r.effectId, (~ok).sum(), np.abs(g - ref).max())
                    shader_checked += 1
                dead = np.all(want[:, 0:4] == 0.0, axis=1)
                assert np.array_equal(np.all(g[:, 0:4] == 0.0, axis=1), dead)
                dead_quads += int(dead.sum()) // 4
                v0 += 4 * r.numParticles
        assert dead_quads > 0, "the scene should contain particles in their last step (collapsed quads)"
        # oracle/_ref travels with the snapshot: on the GPU box the executed shader must have been the checker
        assert shader_checked == 36 or not G.have()
    finally:
        ctx.close()


def test_shade_vertices_capacity_cut_and_errors():
    import torch
    P = _particles()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 12, max_particles_per_frame=1 << 20)
    try:
        s = ctx.create_system((1, 2, 3, 4))
        e = s.add_effect(S.comp_c1(burst=500))
        s.update([e], abi.make_frame_args(0.02, camera=CAMERA))
        ra = make_render_args(CAMERA)
        for k in range(4):
            for pl in range(4):
                ra.frustum.side[k][pl] = 1.0 if k == 3 else 0.0
                ra.frustum.caps[k][pl] = 1.0 if k == 3 else 0.0
        recs, _ = s.render([e], ra)
        assert len(recs) == 1 and recs[0].numParticles >= 500
        out = torch.full((4 * 100 + 4, 16), -1.0, dtype=torch.float32, device="cuda")
        n = s.shade_vertices(recs, out.data_ptr(), 4 * 100 + 3)   # room for 100 quads and a bit: cut at a quad boundary
        ctx.synchronize(); torch.cuda.synchronize()
        assert n == 400
        assert np.all(out[400:].cpu().numpy() == -1.0)
        assert s.shade_vertices([], out.data_ptr(), 400) == 0
        bad = abi.DrawRecord(); C.memmove(C.byref(bad), C.byref(recs[0]), C.sizeof(bad)); bad.typeId = 99
        with pytest.raises(Exception):
            s.shade_vertices([bad], out.data_ptr(), 400)
    finally:
        ctx.close()
