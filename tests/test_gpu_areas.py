"""GPU: sprQueryEffects (k_query_frustum) against the golden vectors of the reference's
cl::AreaSystem and against the oracle on seeded scenes of adds, moves, removes and recycled ids; then the lists it
returns drive update + render exactly like the oracle's particle system given the same lists."""
import hashlib
import json
import os

import numpy as np
import pytest

import scenarios as S
from oracle import areas as A
from oracle import orc
from spear_b200 import abi
from tests_render_args import area_test_frusta, make_render_args

pytestmark = pytest.mark.gpu

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "areas.json")))


def _particles():
    from spear_b200 import particles
    return particles


def _comp(radius):
    # an effect that never spawns: only its sphere area matters here
    return abi.make_component(timeStep=1 / 60, burstAmount=0, spawnTime=1e9, updateRadius=radius)


def run_scene_product(system, ops, frusta, other_system=None):
    """The scene through the C ABI: addEffect / updateTransform / remove / sprQueryEffects. Effect ids are recycled
    LIFO by the product, so results are mapped back to the scene's user ids."""
    comps, effect_of, user_of, tr_of, out = {}, {}, {}, {}, []
    for i, op in enumerate(ops):
        if op[0] == "add":
            c = comps.setdefault(op[3], _comp(op[3]))
            tr = abi.make_transform(op[2])
            e = system.add_effect(c, tr)
            effect_of[op[1]], user_of[e], tr_of[op[1]] = e, op[1], tr
            if other_system is not None and i % 5 == 0:   # another system of the same context never leaks into the result
                other_system.add_effect(c, tr)
        elif op[0] == "move":
            tr = abi.make_transform(op[2])
            system.update_transform(effect_of[op[1]], S.transform_update(tr_of[op[1]], tr))
            tr_of[op[1]] = tr
        elif op[0] == "remove":
            e = effect_of.pop(op[1])
            system.remove(e)
            del user_of[e]
        elif op[0] == "query":
            ids, count = system.query_effects(frusta[op[1]])
            assert count == len(ids)
            assert np.all(ids[1:] > ids[:-1])                      # ascending effect ids, no duplicates
            out.append(np.sort(np.array([user_of[int(e)] for e in ids], dtype=np.uint32)))
    return out


@pytest.mark.parametrize("scene", GOLDEN["scenes"], ids=lambda s: "n%d" % s["n"])
def test_query_matches_golden_vectors(scene):
    P = _particles()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 12)
    try:
        other = ctx.create_system((9, 9, 9, 9))
        res = run_scene_product(ctx.create_system(), A.random_scene(scene["seed"], scene["n"]), area_test_frusta(), other)
        assert [len(r) for r in res] == scene["counts"]
        assert [hashlib.sha256(r.astype("<u4").tobytes()).hexdigest() for r in res] == scene["sha256"]
    finally:
        ctx.close()


@pytest.mark.parametrize("seed,n,in_slab", [(41, 0, True), (42, 7, True), (43, 1500, True), (44, 1500, False), (45, 70000, True)])
def test_query_matches_oracle(seed, n, in_slab):
    P = _particles()
    ops = A.random_scene(seed, n, rounds=3, in_slab=in_slab)
    frusta = area_test_frusta()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 12)
    try:
        got = run_scene_product(ctx.create_system(), ops, frusta)
    finally:
        ctx.close()
    want = A.run_scene_port(ops, frusta)
    assert len(got) == len(want)
    for q, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), "query %d" % q
    if in_slab and A.have_reference() and n <= 2000:
        for q, (a, b) in enumerate(zip(got, A.run_scene_reference(ops, frusta))):
            assert np.array_equal(a, b), "reference query %d" % q


def test_query_capacity_and_empty_system():
    P = _particles()
    import ctypes as C
    ctx = P.ParticleContext(initial_slot_capacity=1 << 12)
    try:
        s = ctx.create_system()
        fr = make_render_args().frustum
        ids, count = s.query_effects(fr)
        assert count == 0 and len(ids) == 0
        c = _comp(1.0)
        for i in range(10):
            s.add_effect(c, abi.make_transform((14.0 + 0.1 * i, 0.0, 14.0)))
        ids, count = s.query_effects(fr, capacity=4)            # the full count comes back, the first `capacity` ids are written
        assert count == 10 and ids.tolist() == [0, 1, 2, 3]
        assert P.lib.sprQueryEffects(s.h, None, None, 0, None) != 0   # argument errors are reported, not crashed on
    finally:
        ctx.close()


def test_query_lists_drive_update_and_render_like_the_oracle(oracle_port):
    """The whole frame with device-made lists: active = visible = sprQueryEffects(frustum); the oracle's particle
    system is given the same lists, and every snapshot and draw record must match bit for bit."""
    P = _particles()
    ra = make_render_args()
    comps = [abi.make_component(timeStep=1 / 60, burstAmount=50 + 30 * k, spawnTime=0.004, spawnTimeVariance=0.002, lifeTime=0.4,
                                lifeTimeVariance=0.2, drag=0.3, gravity=(0, -9.81, 0), updateRadius=0.5 + k,
                                emitPosition=dict(boxExtent=(1, 0.5, 1)), emitVelocity=dict(sphere=dict(minRadius=1, maxRadius=3)))
             for k in range(3)]
    ctx = P.ParticleContext(initial_slot_capacity=1 << 14)
    try:
        seed = (1, 77, 3, 4)
        s, o = ctx.create_system(seed), orc.OracleSystem(oracle_port, seed)
        rng = np.random.default_rng(5)
        pos, radius = {}, {}
        for i in range(40):
            p = (float(rng.uniform(-40, 60)), float(rng.uniform(0, 4)), float(rng.uniform(-40, 60)))
            tr = abi.make_transform(p)
            e = s.add_effect(comps[i % 3], tr)
            assert e == o.add_effect(comps[i % 3], tr)
            pos[e], radius[e] = tr, comps[i % 3].updateRadius
        seen = set()
        for f in range(30):
            if f % 5 == 4:                                       # move a few effects across the frustum boundary
                for e in rng.choice(sorted(pos), 6, replace=False):
                    p = (float(rng.uniform(-40, 60)), float(rng.uniform(0, 4)), float(rng.uniform(-40, 60)))
                    tr = abi.make_transform(p)
                    u = S.transform_update(pos[int(e)], tr)
                    s.update_transform(int(e), u); o.update_transform(int(e), u)
                    pos[int(e)] = tr
            ids, _ = s.query_effects(ra.frustum)
            live = sorted(pos)
            want = A.port_query(ra.frustum, live, [(pos[e].position.x, pos[e].position.y, pos[e].position.z) for e in live],
                                [radius[e] for e in live])
            assert ids.tolist() == want.tolist()
            seen.update(ids.tolist())
            fa = abi.make_frame_args(1 / 60, game_time=f / 60, frame_index=f, camera=(16.0, 20.0, -30.0))
            s.update(ids, fa); o.update(ids, fa)
            r1, n1 = s.render(ids, ra)
            r2, n2 = o.render(ids, ra)
            assert n1 == n2 and [r.effectId for r in r1] == [r.effectId for r in r2]
        assert 0 < len(seen) < 40
        all_ids = sorted(pos)
        S.assert_snapshots_equal(S.snapshot(o, all_ids), S.snapshot(s, all_ids), "query-driven frames")
    finally:
        ctx.close()
