#!/usr/bin/env python3
"""Generate tests/golden/*.json from the REFERENCE ITSELF (oracle/_ref, built from
/root/reference/src by oracle/build_ref.py). Run in the build container only:

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these files are
what pins the plain-C oracle (tests/test_oracle.py) and, on the GPU box where the reference
tree does not exist, the CUDA path (tests/test_gpu_golden.py).

Contents:
  scenarios.json   per scenario, per frame: RNG state, SprEffectInfo bits (counts, timers, bounds),
                   sha256 of freeIndices and of the vertex stream -- from libspear_ref.so (SSE Float4);
                   the gravity-point scenario is taken from libspear_ref_scalar.so (IEEE 1/sqrt),
                   the only case where the two reference builds differ.
  c1_full.json     BASELINE configs[0]: 10k burst, 600 fixed steps: end-state counts, RNG state,
                   cumulative particle-steps, stream hash (known answers of SURVEY.md Appendix C).
  types.json       LUT bytes (gradient alpha masked: uninitialised in the reference) and type UBO bytes.
  rng.json         sf::Random known answers.
  render.json      renderMain draw records of a 20-effect scene over 6 frames.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import scenarios as S  # noqa: E402
from oracle import build_ref, orc  # noqa: E402
from spear_b200 import abi  # noqa: E402
from tests_render_args import make_render_args  # noqa: E402


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=1 if len(json.dumps(obj)) < 20000 else None, sort_keys=True, separators=(",", ":") if len(json.dumps(obj)) >= 20000 else None)
    print("wrote", name)


def render_scene(system):
    comps = [S.comp_c2(k, burst=250, churn=True) for k in range(4)] + [S.comp_local_space()]
    ids = []
    for i in range(20):
        tr = abi.make_transform((4.0 * (i % 5), 0.0, 4.0 * (i // 5) - (40.0 if i == 19 else 0.0)), (0, 0.38268343, 0, 0.92387953), 1.0)
        ids.append(system.add_effect(comps[i % len(comps)], tr))
    ra = make_render_args()
    frames = []
    for frame in range(6):
        fa = abi.make_frame_args(0.0171, game_time=frame * 0.0171, frame_index=frame, camera=S.CAMERA)
        system.update(ids, fa)
        vis = ids if frame % 2 == 0 else ids[::-1]
        recs, nsorted = system.render(vis, ra, capacity=64) if hasattr(system, "kind") else system.render(vis, ra)
        frames.append(dict(numSorted=nsorted, records=[
            dict(effectId=r.effectId, typeId=r.typeId, numParticles=r.numParticles, typeUboChanged=r.typeUboChanged,
                 ubo=hashlib.sha256(bytes(r.instanceUbo)[:152]).hexdigest()[:16]) for r in recs]))
    return frames


def render_capture_scene(o):
    """Row N1: what the renderer side of cl::ParticleSystem hands to sokol over 40 frames -- the bytes appended to the
    16-deep vertex-buffer ring (ParticleSystem.cpp:861-867: re-upload when the cached copy is 16 frames old, per-frame
    budget of 4096 particles with the early `return`), the bindings of every draw (:892-899), the spline atlas texels
    (:421-449, with a type id released and taken over by another descriptor) and the sphere areas kept current by
    updateTransform (:782-783). `o` is an OracleSystem of kind "reference" or "adapter"."""
    comps = [S.comp_c2(k, burst=420, churn=True) for k in range(4)] + [S.comp_local_space()]
    ids = []
    for i in range(12):
        tr = abi.make_transform((4.0 * (i % 4), 0.0, 4.0 * (i // 4)), (0, 0.38268343, 0, 0.92387953), 1.0)
        ids.append(o.add_effect(comps[i % len(comps)], tr))
    extra = abi.make_component(timeStep=1 / 60, burstAmount=300, spawnTime=0.01, lifeTime=0.5, renderOrder=2.0,
                               emitPosition=dict(boxExtent=(1, 1, 1)), scaleSpline=[(0, 1), (1, 0)], gradient=dict(points=[(0, (0.2, 0.4, 1.0))]))
    other = abi.make_component(timeStep=1 / 60, burstAmount=200, spawnTime=1e9, lifeTime=3.0, renderOrder=-1.0,
                               emitPosition=dict(sphere=dict(minRadius=0.1, maxRadius=1.0)), alphaSpline=[(0, 0.3), (1, 0.9)])
    ra = make_render_args()
    frames = []
    where = {}
    for frame in range(40):
        if frame == 5:
            ids.append(o.add_effect(extra, abi.make_transform((1, 1, 1))))
        if frame == 12:                       # the only effect of `extra` goes away: its type id is free again ...
            o.remove(ids.pop())
        if frame == 14:                       # ... and `other` takes it over: its LUT rows replace the cell's texels
            ids.append(o.add_effect(other, abi.make_transform((2, 0, 2))))
        if frame in (3, 9, 21):               # moving emitters: the area spheres follow
            for e in ids[:4]:
                tr = abi.make_transform((0.5 * frame + e, 0.25 * e, -1.0 * e), (0, 0, 0, 1), 1.0 + 0.1 * e)
                o.update_transform(e, S.transform_update(where.get(e, abi.make_transform()), tr))
                where[e] = tr
        fa = abi.make_frame_args(0.0171, game_time=frame * 0.0171, frame_index=frame, camera=S.CAMERA)
        # from frame 18 on every other effect stops being simulated: its cached upload is drawn from the ring slot it
        # went into and is appended again once it is 16 frames old
        o.update(ids if frame < 18 else ids[::2], fa)
        vis = ids if frame % 2 == 0 else ids[::-1]
        recs, nsorted = o.render(vis, ra, capacity=64)
        draws, appends = o.render_capture()
        frames.append(dict(numSorted=nsorted, draws=draws, appends=appends, records=[
            dict(effectId=r.effectId, typeId=r.typeId, numParticles=r.numParticles, typeUboChanged=r.typeUboChanged,
                 ubo=hashlib.sha256(bytes(r.instanceUbo)[:152]).hexdigest()[:16]) for r in recs]))
    atlas = o.read_atlas()
    assert atlas is not None
    atlas[1::2, :, 3] = 0                     # gradient rows: the alpha byte is uninitialised stack in the reference (:326, :366-400)
    used = sorted(set(int(y) // 2 + 256 * (int(x) // 64) for y, x in zip(*np.nonzero(atlas.any(axis=2)))))
    return dict(frames=frames, atlas=hashlib.sha256(atlas.tobytes()).hexdigest(), atlasCells=used,
                spheres=[[S.f32_bits(v) for v in s] for s in o.area_spheres()])


def main():
    build_ref.build(verbose=False)
    out = {}
    for sc in S.scenarios():
        kind = "reference" if sc.float_exact else "reference_scalar"
        snaps = S.run_scenario(orc.OracleSystem(kind, sc.seed), sc)
        out[sc.name] = dict(kind=kind, frames=[S.snapshot_digest(s) for s in snaps])
    dump("scenarios.json", out)

    o = orc.OracleSystem("reference", (1, 2, 3, 4))
    e = o.add_effect(S.comp_c1(), abi.make_transform((1, 2, 3)))
    dt = o.effect_info(e).timeStep
    total = 0
    for i in range(600):
        o.update([e], abi.make_frame_args(dt, game_time=i * dt, frame_index=i))
        total += o.effect_info(e).aliveCount
    info = o.effect_info(e)
    dump("c1_full.json", dict(slotCount=info.slotCount, aliveCount=info.aliveCount, freeCount=info.freeCount,
                              particleSteps=total, rng="%016x" % o.rng_state()[0], timeStepBits=S.f32_bits(info.timeStep),
                              stream=hashlib.sha256(o.read_stream(e).tobytes()).hexdigest(),
                              free=hashlib.sha256(o.read_free(e).tobytes()).hexdigest()))

    types = {}
    comps = dict(c1=S.comp_c1(), fountain=S.comp_fountain(), c2_0=S.comp_c2(0), c2_1=S.comp_c2(1), c2_2=S.comp_c2(2), c2_3=S.comp_c2(3),
                 defaults=abi.make_component())
    o = orc.OracleSystem("reference")
    for name, comp in comps.items():
        t = o.reserve_type(comp)
        lut = o.type_lut(t).copy()
        lut[256 + 3::4] = 0
        types[name] = dict(typeId=t, lut=lut.tobytes().hex(), ubo=bytes(o.type_ubo(t)).hex())
    dump("types.json", types)

    # sf::Random(2, 581271): states and outputs, nextFloat bits
    import ctypes as C
    class R:  # the reference RNG is only reachable through a system; restate and cross-check below
        def __init__(self, s, i): self.s, self.i = s, i | 1
        def u32(self):
            old = self.s
            self.s = (old * 6364136223846793005 + self.i) & (2**64 - 1)
            x = (((old >> 18) ^ old) >> 27) & 0xffffffff
            rot = old >> 59
            return ((x >> rot) | (x << ((-rot) & 31))) & 0xffffffff
    r = R(2, 581271)
    rows = []
    for _ in range(6):
        st = r.s
        u = r.u32()
        rows.append(dict(state="%016x" % st, u32="%08x" % u, floatBits="%08x" % S.f32_bits(float(np.float32(np.float64(u) * 2.3283064365386963e-10)))))
    # cross-check against the reference: after the burst6 scenario the system RNG has advanced exactly 50 draws
    r2 = R(2, 581271)
    for _ in range(50):
        r2.u32()
    sc = [s for s in S.scenarios() if s.name == "burst6"][0]
    o = orc.OracleSystem("reference", sc.seed)
    e = o.add_effect(*sc.effects[0])
    o.update([e], abi.make_frame_args(o.effect_info(e).timeStep))
    assert o.rng_state()[0] == r2.s, "PCG restatement disagrees with the reference"
    dump("rng.json", dict(rows=rows, after50="%016x" % r2.s))

    dump("render.json", render_scene(orc.OracleSystem("reference", (1, 5, 3, 4))))
    dump("render_capture.json", render_capture_scene(orc.OracleSystem("reference", (1, 6, 3, 4))))


if __name__ == "__main__":
    main()
