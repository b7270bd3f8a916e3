"""More GPU parity: large multi-tile effects, per-particle depth sort, many systems in one
launch set, renderMain draw records, type LUT/UBO, and full-size property checks."""
import ctypes as C
import os

import numpy as np
import pytest

import scenarios as S
from oracle import orc
from spear_b200 import abi

pytestmark = pytest.mark.gpu

CAMERA = (16.0, 20.0, -30.0)


def _particles():
    from spear_b200 import particles
    return particles


sort_key, expected_sorted = S.sort_key, S.expected_sorted


def comp_big(burst):
    return abi.make_component(
        timeStep=1 / 60, burstAmount=burst, spawnTime=0.0001, lifeTime=0.05, lifeTimeVariance=0.1, drag=0.2,
        gravity=(0, -9.81, 0), emitPosition=dict(sphere=dict(minRadius=0.5, maxRadius=4.0)),
        emitVelocity=dict(boxExtent=(2, 2, 2), sphere=dict(minRadius=0.1, maxRadius=1.0)))


def test_large_effect_multi_tile(oracle_port):
    """200k-particle burst: wide burst kernel, chained look-back across ~200 tiles, deaths and
    re-use of freed slots across tiles."""
    P = _particles()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 16)
    try:
        sc = S.Scenario("big", (3, 1234, 5, 6), [(comp_big(200_000), abi.make_transform((1, 2, 3)))], frames=12)
        ref = S.run_scenario(orc.OracleSystem(oracle_port, sc.seed), sc)
        got = S.run_scenario(ctx.create_system(sc.seed), sc)
        for f, (a, b) in enumerate(zip(ref, got)):
            S.assert_snapshots_equal(a, b, "big frame %d" % f)
        assert ref[0][0]["info"]["aliveCount"] == 200_020
        assert ref[-1][0]["info"]["freeCount"] > 1000
    finally:
        ctx.close()


@pytest.mark.parametrize("name", ["c1_small", "fountain", "multi", "churn_c5", "remove_reuse", "gravity_points"])
def test_sorted_stream_is_stable_depth_sort_of_reference_stream(oracle_port, name):
    P = _particles()
    sc = [s for s in S.scenarios() if s.name == name][0]
    ctx = P.ParticleContext(sort_particles=True, initial_slot_capacity=1 << 16)
    try:
        ref = S.run_scenario(orc.OracleSystem(oracle_port, sc.seed), sc)
        sys_ = ctx.create_system(sc.seed)
        frames = []

        def on_frame(frame, snap):
            ids = [e for e in snap if e != "rng"]
            frames.append({e: sys_.read_stream_slots(e).copy() for e in ids})

        got = S.run_scenario(sys_, sc, on_frame=on_frame)
        for f, (a, b) in enumerate(zip(ref, got)):
            assert a["rng"] == b["rng"]
            for e in a:
                if e == "rng":
                    continue
                assert a[e]["info"] == b[e]["info"], "%s frame %d effect %d" % (name, f, e)
                assert np.array_equal(a[e]["free"], b[e]["free"])
                exp, order = expected_sorted(a[e]["stream"])
                assert exp.tobytes() == b[e]["stream"].tobytes(), "%s frame %d effect %d sorted stream" % (name, f, e)
    finally:
        ctx.close()


def test_sorted_large_effect(oracle_port):
    P = _particles()
    ctx = P.ParticleContext(sort_particles=True, initial_slot_capacity=1 << 16)
    try:
        sc = S.Scenario("bigsort", (3, 77, 5, 6), [(comp_big(70_000), abi.make_transform((1, 2, 3)))], frames=6)
        ref = S.run_scenario(orc.OracleSystem(oracle_port, sc.seed), sc)
        got = S.run_scenario(ctx.create_system(sc.seed), sc)
        for f, (a, b) in enumerate(zip(ref, got)):
            assert a[0]["info"] == b[0]["info"]
            exp, _ = expected_sorted(a[0]["stream"])
            assert exp.tobytes() == b[0]["stream"].tobytes(), "frame %d" % f
    finally:
        ctx.close()


def test_sorted_with_a_moving_camera(oracle_port):
    """The depth keys follow the camera of every frame: a camera that drifts, sits inside the cloud, is far away (the
    keys' top digits become constant: pass skipping) and jumps between frames, over a tile-path effect (70k slots), a
    one-CTA effect (6000) and a one-granule effect (40) of one system. Every frame's stream is the stable depth sort of
    the oracle's stream under THAT frame's camera."""
    P = _particles()
    ctx = P.ParticleContext(sort_particles=True, initial_slot_capacity=1 << 16)
    try:
        seed = (9, 8, 7, 6)
        effects = [(comp_big(70_000), abi.make_transform((1, 2, 3))), (comp_big(6_000), abi.make_transform((-4, 1, 2))),
                   (S.comp_c5(), abi.make_transform((2, 0, -1)))]
        o = orc.OracleSystem(oracle_port, seed)
        s = ctx.create_system(seed)
        ids = [s.add_effect(c, t) for c, t in effects]
        assert ids == [o.add_effect(c, t) for c, t in effects]
        cameras = [(16.0, 20.0, -30.0), (15.5, 19.0, -29.0), (1.0, 2.0, 3.0), (1.5, 2.5, 2.0), (4000.0, 3000.0, -9000.0),
                   (-60.0, 5.0, 12.0), (16.0, 20.0, -30.0), (0.0, 0.0, 0.0), (1e6, 1e6, 1e6), (2.0, 3.0, 4.0)]
        dt = 1 / 60
        for f, cam in enumerate(cameras):
            fa = abi.make_frame_args(dt, game_time=f * dt, frame_index=f, camera=cam)
            s.update(ids, fa); o.update(ids, fa)
            a, b = S.snapshot(o, ids), S.snapshot(s, ids)
            for e in ids:
                assert a[e]["info"] == b[e]["info"], "frame %d effect %d" % (f, e)
                exp, order = expected_sorted(a[e]["stream"], cam)
                assert exp.tobytes() == b[e]["stream"].tobytes(), "frame %d effect %d sorted stream under camera %r" % (f, e, cam)
    finally:
        ctx.close()


def test_batched_update_sorts_every_system_towards_its_own_camera(oracle_port):
    """sprUpdateParticlesBatch with one FrameArgs per system (argsStride != 0): dt, game time and the camera of the depth
    keys are per system -- the batch equals the per-system calls. Three systems with the same effects, different cameras
    and time steps; one of them is also driven alone in a second context and must give the same bytes."""
    P = _particles()
    ctx = P.ParticleContext(sort_particles=True, initial_slot_capacity=1 << 16)
    solo = P.ParticleContext(sort_particles=True, initial_slot_capacity=1 << 16)
    try:
        effects = [(comp_big(30_000), abi.make_transform((1, 2, 3))), (comp_big(3_000), abi.make_transform((-4, 1, 2))),
                   (S.comp_c5(), abi.make_transform((2, 0, -1)))]
        seeds = [(1, 2, 3, 4), (5, 6, 7, 8), (9, 10, 11, 12)]
        cams = [(16.0, 20.0, -30.0), (-3.0, 1.0, 2.0), (500.0, -40.0, 90.0)]
        dts = [1 / 60, 1 / 45, 1 / 75]
        oracles, systems, ids = [], [], []
        for sd in seeds:
            o = orc.OracleSystem(oracle_port, sd); s = ctx.create_system(sd)
            my = [s.add_effect(c, t) for c, t in effects]
            assert my == [o.add_effect(c, t) for c, t in effects]
            oracles.append(o); systems.append(s); ids.append(my)
        s1 = solo.create_system(seeds[1])
        ids1 = [s1.add_effect(c, t) for c, t in effects]
        sorted_under = {}   # (system, effect) -> (simSteps, camera of the last frame that simulated the effect: a frame
                            # whose dt does not reach the effect's time step leaves its stream as it is)
        for f in range(8):
            fas = [abi.make_frame_args(dts[i], game_time=f * dts[i], frame_index=f, camera=(cams[i][0] + f, cams[i][1], cams[i][2])) for i in range(3)]
            ctx.update_batch(systems, ids, fas)
            s1.update(ids1, fas[1])
            for i in range(3):
                oracles[i].update(ids[i], fas[i])
                a, b = S.snapshot(oracles[i], ids[i]), S.snapshot(systems[i], ids[i])
                for e in ids[i]:
                    assert a[e]["info"] == b[e]["info"], "frame %d system %d effect %d" % (f, i, e)
                    steps = a[e]["info"]["simSteps"]
                    if sorted_under.get((i, e), (0, None))[0] != steps:
                        sorted_under[(i, e)] = (steps, (cams[i][0] + f, cams[i][1], cams[i][2]))
                    if steps == 0:
                        continue
                    exp, _ = expected_sorted(a[e]["stream"], sorted_under[(i, e)][1])
                    assert exp.tobytes() == b[e]["stream"].tobytes(), "frame %d system %d effect %d" % (f, i, e)
            c = S.snapshot(s1, ids1)
            b = S.snapshot(systems[1], ids[1])
            for e in ids1:
                assert c[e]["stream"].tobytes() == b[e]["stream"].tobytes(), "frame %d: batch differs from the per-system call" % f
    finally:
        ctx.close(); solo.close()


@pytest.mark.parametrize("sort", [False, True], ids=["stream", "sorted"])
def test_effects_above_4gb_of_vertex_stream(oracle_port, sort):
    """Maximum sizes: an idle 40M-particle effect is registered first, so every effect of the scenario lives
    above slot 2^26 -- state float4 indices above 2^27, vertex byte offsets above 2^33, i.e. every 32-bit
    index product in the kernels would overflow. The scenario must still match the oracle bit for bit."""
    P = _particles()
    ctx = P.ParticleContext(sort_particles=sort, initial_slot_capacity=1 << 12, max_particles_per_frame=1 << 30)
    try:
        sc = [s for s in S.scenarios() if s.name == "multi"][0]
        ballast = abi.make_component(burstAmount=40_000_000, lifeTime=10.0)
        o = orc.OracleSystem(oracle_port, sc.seed)
        s = ctx.create_system(sc.seed)
        assert o.add_effect(ballast) == s.add_effect(ballast) == 0      # never in an active list: it only occupies the range
        ref = S.run_scenario(o, sc)
        got = S.run_scenario(s, sc)
        first = min(e for e in got[0] if e != "rng")
        assert s.device_view(first).vertexByteOffset >= 1 << 33
        for f, (a, b) in enumerate(zip(ref, got)):
            if not sort:
                S.assert_snapshots_equal(a, b, "high offsets frame %d" % f)
                continue
            assert a["rng"] == b["rng"]
            for e in a:
                if e == "rng":
                    continue
                assert a[e]["info"] == b[e]["info"], "frame %d effect %d" % (f, e)
                assert np.array_equal(a[e]["free"], b[e]["free"])
                exp, _ = expected_sorted(a[e]["stream"])
                assert exp.tobytes() == b[e]["stream"].tobytes(), "frame %d effect %d sorted stream" % (f, e)
        # the draw records carry the 64-bit offsets
        from tests_render_args import make_render_args
        ids = [e for e in got[-1] if e != "rng"]
        recs, _ = s.render(ids, make_render_args())
        assert recs and all(r.vertexByteOffset >= 1 << 33 for r in recs)
    finally:
        ctx.close()


def test_sorted_pass_skipping_shapes(oracle_port):
    """The radix sort runs on key - keyMin and skips passes whose digit is constant: effects built so that 1, 2, 3 and
    all 4 passes execute (odd counts leave the result in the other ping-pong buffer), checked against the stable
    sort of the oracle's stream."""
    P = _particles()
    ctx = P.ParticleContext(sort_particles=True, initial_slot_capacity=1 << 12)
    try:
        still = dict(drag=0.0, gravity=(0, 0, 0), lifeTime=1e9, spawnTime=1e9, timeStep=1 / 60)
        comps = [
            # every particle at the same point: all keys equal, only pass 0 runs (ties -> slot order)
            abi.make_component(burstAmount=3000, **still),
            # a thin shell far away: squared distances differ in the low 16 bits only -> 2 passes
            abi.make_component(burstAmount=5000, emitPosition=dict(boxExtent=(0.002, 0.002, 0.002)), **still),
            # a box of a few metres seen from 40 m: < 2^24 distinct keys -> 3 passes
            abi.make_component(burstAmount=20000, emitPosition=dict(boxExtent=(4, 2, 4)), emitVelocity=dict(boxExtent=(1, 1, 1)), **still),
            # a cloud around the camera: squared distances from ~0 to ~10^5, every exponent -> 4 passes
            abi.make_component(burstAmount=20000, emitPosition=dict(boxExtent=(400, 300, 400)), **still),
        ]
        trs = [abi.make_transform((1, 2, 3)), abi.make_transform((300, 2, 3)), abi.make_transform((1, 2, 3)), abi.make_transform(CAMERA)]
        sseed = (1, 55, 3, 4)
        s, o = ctx.create_system(sseed), orc.OracleSystem(oracle_port, sseed)
        ids = [s.add_effect(c, t) for c, t in zip(comps, trs)]
        assert ids == [o.add_effect(c, t) for c, t in zip(comps, trs)]
        for f in range(4):
            fa = abi.make_frame_args(1 / 55, game_time=f / 55, frame_index=f, camera=CAMERA)
            s.update(ids, fa); o.update(ids, fa)
            a, b = S.snapshot(o, ids), S.snapshot(s, ids)
            assert a["rng"] == b["rng"]
            for e in ids:
                assert a[e]["info"] == b[e]["info"], "frame %d effect %d" % (f, e)
                exp, _ = expected_sorted(a[e]["stream"])
                assert exp.tobytes() == b[e]["stream"].tobytes(), "frame %d effect %d sorted stream" % (f, e)
                slots = s.read_stream_slots(e)
                assert len(slots) == a[e]["info"]["aliveCount"] and len(set(slots.tolist())) == len(slots)
    finally:
        ctx.close()


def test_sorted_small_effects_one_cta_path(oracle_port):
    """Effects of at most 18432 slots are sorted whole by one CTA (k_sort_small): 1 to 18 keys per thread, 1 to 4
    executed passes, next to a tile-path effect in the same frame, and an effect that grows across the limit while it
    is simulated (one-CTA path first, tile path later) -- all against the stable sort of the oracle's stream."""
    P = _particles()
    ctx = P.ParticleContext(sort_particles=True, initial_slot_capacity=1 << 12)
    try:
        still = dict(drag=0.0, gravity=(0, 0, 0), lifeTime=1e9, spawnTime=1e9, timeStep=1 / 60)
        comps = [
            abi.make_component(burstAmount=700, **still),                                                               # 1 pass, 1 key per thread
            abi.make_component(burstAmount=1025, emitPosition=dict(boxExtent=(0.002, 0.002, 0.002)), **still),           # 2 passes, 2 keys per thread
            abi.make_component(burstAmount=18300, emitPosition=dict(boxExtent=(4, 2, 4)), emitVelocity=dict(boxExtent=(1, 1, 1)), **still),  # 3 passes, 18 keys per thread
            abi.make_component(burstAmount=17000, emitPosition=dict(boxExtent=(400, 300, 400)), **still),                # 4 passes
            abi.make_component(burstAmount=30000, emitPosition=dict(boxExtent=(4, 2, 4)), **still),                      # tile path, same frame
            # grows by 20 particles per step from 18200: crosses the 18432-slot limit after a dozen steps; short lives
            # keep holes in the slot range (the first pass compacts under the alive ballots)
            abi.make_component(burstAmount=18200, spawnTime=1e-5, lifeTime=1e9, drag=0.1, gravity=(0, -9.81, 0), timeStep=1 / 60,
                               emitPosition=dict(boxExtent=(3, 1, 3)), emitVelocity=dict(sphere=dict(minRadius=0.5, maxRadius=2.0))),
            abi.make_component(burstAmount=4000, spawnTime=0.002, lifeTime=0.1, lifeTimeVariance=0.2, drag=0.1, gravity=(0, -9.81, 0), timeStep=1 / 60,
                               emitPosition=dict(boxExtent=(3, 1, 3)), emitVelocity=dict(sphere=dict(minRadius=0.5, maxRadius=2.0))),
            # a ball seen from outside: the depth histogram is a hump (full middle digits, thin ends), 3 significant bytes
            abi.make_component(burstAmount=15000, emitPosition=dict(sphere=dict(minRadius=0.0, maxRadius=6.0)), **still),
            # a few hundred keys spread over 3 bytes: most digit bins empty in every pass
            abi.make_component(burstAmount=300, emitPosition=dict(boxExtent=(4, 2, 4)), **still),
        ]
        trs = [abi.make_transform((1, 2, 3)), abi.make_transform((300, 2, 3)), abi.make_transform((1, 2, 3)), abi.make_transform(CAMERA),
               abi.make_transform((5, 2, 3)), abi.make_transform((-4, 1, 6)), abi.make_transform((9, 1, -6)), abi.make_transform((-20, 3, 25)),
               abi.make_transform((30, 0, 10))]
        sseed = (1, 91, 3, 4)
        s, o = ctx.create_system(sseed), orc.OracleSystem(oracle_port, sseed)
        ids = [s.add_effect(c, t) for c, t in zip(comps, trs)]
        assert ids == [o.add_effect(c, t) for c, t in zip(comps, trs)]
        crossed = False
        for f in range(24):
            fa = abi.make_frame_args(1 / 55, game_time=f / 55, frame_index=f, camera=CAMERA)
            s.update(ids, fa); o.update(ids, fa)
            if f % 3 and f < 20:
                continue
            a, b = S.snapshot(o, ids), S.snapshot(s, ids)
            assert a["rng"] == b["rng"]
            for e in ids:
                assert a[e]["info"] == b[e]["info"], "frame %d effect %d" % (f, e)
                exp, _ = expected_sorted(a[e]["stream"])
                assert exp.tobytes() == b[e]["stream"].tobytes(), "frame %d effect %d sorted stream" % (f, e)
                slots = s.read_stream_slots(e)
                assert len(slots) == a[e]["info"]["aliveCount"] and len(set(slots.tolist())) == len(slots)
            crossed = crossed or a[ids[5]]["info"]["slotCount"] > 18432
        assert crossed
    finally:
        ctx.close()


def test_large_batch_matches_oracle(oracle_port):
    """A frame of 64 systems x 512 effects through sprUpdateParticlesBatch: mostly short-lived effects whose slot range
    is final (finished in the first phase of the host pass), with growing and bursting effects in between (noted
    there, finished in request order in the second phase): RNG state of every system, and full snapshots of every
    8th, against one oracle instance per system."""
    P = _particles()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 18)
    try:
        comps = [S.comp_c5(), S.comp_fountain(), S.comp_c2(1, burst=1500, churn=True)]
        systems, oracles, ids = [], [], []
        for i in range(64):
            seed = (1, 500 + i, 3, 4)
            s, o = ctx.create_system(seed), orc.OracleSystem(oracle_port, seed)
            my = []
            for e in range(512):
                comp = comps[0] if (e + i) % 37 else comps[1 + (e & 1)]
                tr = abi.make_transform((i * 0.5, (e & 15) * 0.25, e * 0.01))
                e1, e2 = s.add_effect(comp, tr), o.add_effect(comp, tr)
                assert e1 == e2
                my.append(e1)
            systems.append(s); oracles.append(o); ids.append(my)
        for frame in range(5):
            fa = abi.make_frame_args(0.0172, game_time=frame * 0.0172, frame_index=frame)
            ctx.update_batch(systems, ids, fa)
            for o, my in zip(oracles, ids):
                o.update(my, fa)
        for i, (s, o, my) in enumerate(zip(systems, oracles, ids)):
            assert s.rng_state() == o.rng_state(), "system %d rng" % i
            if i % 8 == 0:
                S.assert_snapshots_equal(S.snapshot(o, my), S.snapshot(s, my), "system %d" % i)
    finally:
        ctx.close()


def test_render_batch_matches_per_system_render():
    """sprRenderMainBatch is defined as sprRenderMain per system. 64 systems x 512 one-particle effects is above the
    size where the batch's per-system passes run on several host threads: its draw records must be those of the
    per-system calls on an identically built second context, field for field."""
    from tests_render_args import make_render_args as fixed_camera_args
    P = _particles()
    comp = abi.make_component(timeStep=1 / 60, burstAmount=1, spawnTime=1e9, lifeTime=1e9, updateRadius=1.0)
    comp2 = abi.make_component(timeStep=1 / 60, burstAmount=2, spawnTime=1e9, lifeTime=1e9, renderOrder=3)
    ra = fixed_camera_args()
    for k in range(4):  # an all-inclusive frustum
        for pl in range(4):
            ra.frustum.side[k][pl] = 1.0 if k == 3 else 0.0
            ra.frustum.caps[k][pl] = 1.0 if k == 3 else 0.0
    ctxs = [P.ParticleContext(initial_slot_capacity=1 << 16) for _ in range(2)]
    try:
        built = []
        for ctx in ctxs:
            rng = np.random.default_rng(3)
            systems, ids = [], []
            for i in range(64):
                s = ctx.create_system((1, 2 + i, 3, 4))
                my = [s.add_effect(comp if (e + i) % 5 else comp2, abi.make_transform(tuple(float(x) for x in rng.uniform(-30, 30, 3)))) for e in range(512)]
                systems.append(s); ids.append(my)
            fa = abi.make_frame_args(0.02, game_time=0.0, frame_index=0, camera=CAMERA)
            ctx.update_batch(systems, ids, fa)
            built.append((systems, ids))
        (sa, ia), (sb, ib) = built
        batch = ctxs[0].make_batch(sa, ia)
        for frame in range(2):   # the second frame exercises the upload ring bookkeeping (uploadFrameIndex)
            outs = ctxs[0].render_prepared(batch, ra)
            for i in range(64):
                recs, nsorted = sb[i].render(ib[i], ra)
                assert outs[i].numRecords == len(recs) == 512 and outs[i].numSortedEffects == nsorted
                got = bytes(C.string_at(outs[i].records, C.sizeof(abi.DrawRecord) * len(recs)))
                want = b"".join(bytes(r) for r in recs)
                # vertexByteOffset is a position in each context's own pool; both contexts were built identically
                assert got == want, "system %d frame %d" % (i, frame)
    finally:
        for ctx in ctxs:
            ctx.close()


def test_many_systems_one_launch_set(oracle_port):
    """sprUpdateParticlesBatch: 24 independent systems (own RNG each, seed[1] = 2 + i) advance in one
    set of launches and each matches its own oracle instance."""
    P = _particles()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 16)
    try:
        comps = [S.comp_c2(k, burst=200 + 31 * k, churn=True) for k in range(4)] + [S.comp_c5(), S.comp_fountain()]
        n = 24
        systems, oracles, ids = [], [], []
        for i in range(n):
            seed = (1, 2 + i, 3, 4)
            s, o = ctx.create_system(seed), orc.OracleSystem(oracle_port, seed)
            my = []
            for k in range(3):
                comp = comps[(i + k) % len(comps)]
                tr = abi.make_transform((i, k, 0))
                e1, e2 = s.add_effect(comp, tr), o.add_effect(comp, tr)
                assert e1 == e2
                my.append(e1)
            systems.append(s); oracles.append(o); ids.append(my)
        for frame in range(25):
            fa = abi.make_frame_args(0.0172, game_time=frame * 0.0172, frame_index=frame)
            ctx.update_batch(systems, ids, fa)
            for o, my in zip(oracles, ids):
                o.update(my, fa)
            if frame % 6 == 0 or frame == 24:
                for s, o, my in zip(systems, oracles, ids):
                    S.assert_snapshots_equal(S.snapshot(o, my), S.snapshot(s, my), "batch frame %d" % frame)
    finally:
        ctx.close()


def look_at_perspective(eye, target, fovy=1.0, aspect=16 / 9, near=0.1, far=200.0):
    eye, target = np.asarray(eye, np.float64), np.asarray(target, np.float64)
    f = target - eye; f /= np.linalg.norm(f)
    r = np.cross(f, [0, 1, 0]); r /= np.linalg.norm(r)
    u = np.cross(r, f)
    view = np.eye(4); view[0, :3] = r; view[1, :3] = u; view[2, :3] = -f
    view[:3, 3] = -view[:3, :3] @ eye
    t = 1.0 / np.tan(fovy / 2)
    proj = np.zeros((4, 4)); proj[0, 0] = t / aspect; proj[1, 1] = t
    proj[2, 2] = (far + near) / (near - far); proj[2, 3] = 2 * far * near / (near - far); proj[3, 2] = -1
    return proj, proj @ view


def make_render_args(oracle_kind):
    proj, w2c = look_at_perspective(CAMERA, (6.0, 0.0, 6.0))
    ra = abi.RenderArgs()
    ra.cameraPosition = abi.Vec3(*CAMERA)
    for c in range(4):
        for r in range(4):
            ra.viewToClip.v[c * 4 + r] = float(np.float32(proj[r, c]))
            ra.worldToClip.v[c * 4 + r] = float(np.float32(w2c[r, c]))
    ra.frustum = orc.make_frustum(list(ra.worldToClip.v), 0.0, kind=oracle_kind)
    ra.visFogWorldMad = abi.Vec4(0.1, 0.2, 0.3, 0.4)
    return ra


@pytest.mark.parametrize("budget", [0, 1 << 20])
def test_render_main_draw_records(oracle_port, budget):
    """renderMain (ParticleSystem.cpp:824-903): culling, effect sort, the 4096-particle frame budget,
    type-UBO change flags and the 160-byte instance UBO."""
    P = _particles()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 16, max_particles_per_frame=budget)
    try:
        seed = (1, 5, 3, 4)
        s, o = ctx.create_system(seed), orc.OracleSystem(oracle_port, seed)
        comps = [S.comp_c2(k, burst=250, churn=True) for k in range(4)] + [S.comp_local_space()]
        ids = []
        for i in range(20):
            comp = comps[i % len(comps)]
            tr = abi.make_transform((4.0 * (i % 5), 0.0, 4.0 * (i // 5) - (40.0 if i == 19 else 0.0)), (0, 0.38268343, 0, 0.92387953), 1.0)
            e1, e2 = s.add_effect(comp, tr), o.add_effect(comp, tr)
            assert e1 == e2
            ids.append(e1)
        ra = make_render_args(oracle_port)
        cut = culled = False
        for frame in range(6):
            fa = abi.make_frame_args(0.0171, game_time=frame * 0.0171, frame_index=frame, camera=CAMERA)
            s.update(ids, fa); o.update(ids, fa)
            vis = ids if frame % 2 == 0 else ids[::-1]
            got, got_sorted = s.render(vis, ra)
            exp, exp_sorted = o.render(vis, ra, capacity=64)
            assert got_sorted == exp_sorted
            if budget == 0:
                assert len(got) == len(exp)
                cut = cut or len(got) < got_sorted
                culled = culled or got_sorted < sum(1 for e in ids if s.effect_info(e).aliveCount > 0)
                pairs = zip(got, exp)
            else:
                assert len(got) == got_sorted
                pairs = zip(got[:len(exp)], exp)
            for a, b in pairs:
                assert (a.effectId, a.typeId, a.numParticles, a.typeUboChanged) == (b.effectId, b.typeId, b.numParticles, b.typeUboChanged)
                assert bytes(a.instanceUbo)[:152] == bytes(b.instanceUbo)[:152]
        if budget == 0:
            assert cut and culled  # the 4096-particle budget cut-off and the frustum cull were both exercised
    finally:
        ctx.close()


def test_prepare_for_remove_and_drain(oracle_port):
    P = _particles()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 16)
    try:
        seed = (1, 6, 3, 4)
        s, o = ctx.create_system(seed), orc.OracleSystem(oracle_port, seed)
        comp = S.comp_timed_emitter()
        e = s.add_effect(comp); assert e == o.add_effect(comp)
        dt = s.effect_info(e).timeStep
        for frame in range(60):
            fa = abi.make_frame_args(dt, game_time=frame * dt, frame_index=frame)
            if frame >= 5:
                assert s.prepare_for_remove(e, fa) == o.prepare_for_remove(e, fa), "frame %d" % frame
            s.update([e], fa); o.update([e], fa)
        assert s.effect_info(e).aliveCount == 0
    finally:
        ctx.close()


def test_full_size_properties():
    """BASELINE configs[2] at full size (10M-particle single effect), size-independent properties:
    alive = burst + 1, every particle replicated 4x, slot order ascending /
    depth order descending, checksum of the sorted stream equals the checksum of the unsorted one."""
    P = _particles()
    burst = 10_000_000
    comp = abi.make_component(timeStep=1 / 60, burstAmount=burst, spawnTime=1e9, lifeTime=1e9, drag=0.3,
                              gravity=(0, -9.81, 0), emitPosition=dict(boxExtent=(8, 4, 8)),
                              emitVelocity=dict(sphere=dict(minRadius=1, maxRadius=3)))
    sums = []
    for sort in (False, True):
        ctx = P.ParticleContext(sort_particles=sort, initial_slot_capacity=1 << 24)
        try:
            s = ctx.create_system((1, 2, 3, 4))
            e = s.add_effect(comp, abi.make_transform((1, 2, 3)))
            dt = s.effect_info(e).timeStep
            for frame in range(3):
                s.update([e], abi.make_frame_args(dt, game_time=frame * dt, frame_index=frame, camera=CAMERA))
            info = s.effect_info(e)
            assert info.aliveCount == burst + 1 and info.slotCount == burst + 4 and info.freeCount == 3
            st = s.read_stream(e)
            assert len(st) == 4 * (burst + 1)
            raw = st.view(np.uint32).reshape(-1, 4, 8)
            assert (raw[:, 0] == raw[:, 1]).all() and (raw[:, 0] == raw[:, 2]).all() and (raw[:, 0] == raw[:, 3]).all()
            recs = st[::4]
            slots = s.read_stream_slots(e)
            if sort:
                k = sort_key(recs)
                assert (k[1:] >= k[:-1]).all()
                ties = k[1:] == k[:-1]
                assert (slots[1:][ties] > slots[:-1][ties]).all()
                assert len(np.unique(slots)) == len(slots)
            else:
                assert (slots[1:] > slots[:-1]).all()
            sums.append(int(raw[:, 0].astype(np.uint64).sum()))
        finally:
            ctx.close()
    assert sums[0] == sums[1]


def test_type_lut_and_ubo(oracle_port):
    P = _particles()
    comps = [S.comp_c1(), S.comp_fountain()] + [S.comp_c2(k) for k in range(4)]
    o = orc.OracleSystem(oracle_port)
    for i, comp in enumerate(comps):
        t = o.reserve_type(comp)
        lut, ubo = P.bake_effect_type(comp, t)
        ref = o.type_lut(t)
        assert np.array_equal(lut, ref)
        assert bytes(ubo) == bytes(o.type_ubo(t))


@pytest.mark.parametrize("share", [0, 1], ids=["whole_gpu_lsu_kernel", "one_cta_per_sm_bulk_async_kernel"])
def test_pack_and_fused_expand_from_run_buffers(oracle_port, share):
    """sprRunBufferPack + sprExpandFromPeers with two run buffers of the same GPU standing in for two
    ranks: the assembled stream is run 0 then run 1, each equal to the effects' own vertex streams. At an assembly
    share of one CTA per SM the expansion is the bulk-asynchronous pipeline (cp.async.bulk loads, tensor stores, the
    runs' tails through the lanes), at the default share the load/store kernel."""
    import torch
    P = _particles()
    ctx = P.ParticleContext(sort_particles=True, initial_slot_capacity=1 << 16)
    ctx.set_assembly_share(share)
    try:
        s = ctx.create_system((1, 11, 3, 4))
        comps = [S.comp_c2(k, burst=700 + 100 * k, churn=True) for k in range(4)]
        ids = [s.add_effect(c, abi.make_transform((k, 0, 0))) for k, c in enumerate(comps)]
        for f in range(5):
            s.update(ids, abi.make_frame_args(0.0175, f * 0.0175, f, camera=CAMERA))
        streams = [s.read_stream(e) for e in ids]
        cap = 8192
        bufs = [ctx.run_buffer_create(cap) for _ in range(2)]
        ctx.run_buffer_pack(ctx.make_pack_list([s, s], ids[:2]), bufs[0])
        ctx.run_buffer_pack(ctx.make_pack_list([s, s], ids[2:]), bufs[1])
        total = sum(len(st) for st in streams) // 4
        out = torch.zeros((2 * cap * 4, 8), dtype=torch.float32, device="cuda")
        tot = torch.zeros(1, dtype=torch.int64, device="cuda")
        ctx.expand_from_peers([ctx.run_buffer_ptr(b) for b in bufs], out.data_ptr(), 2 * cap, tot.data_ptr())
        ctx.synchronize(); torch.cuda.synchronize()
        assert int(tot.item()) == total
        got = out[:total * 4].cpu().numpy().tobytes()
        assert got == b"".join(st.tobytes() for st in streams)
    finally:
        ctx.close()


def test_deferred_expansion_assembles_the_same_stream(oracle_port):
    """sprContextSetDeferredExpansion: updates sort but do not write the local vertex stream, the pack calls gather the
    sorted records from the slot state -- the assembled stream (sprRunBufferPack -> sprExpandFromPeers and sprPackRuns ->
    sprExpandRuns) equals the one built from the local streams of an identical context, for one-granule, one-CTA and tile
    path effects (odd and even numbers of executed radix passes), over frames with deaths and spawns; switching it off
    again brings the local stream back with the next update; a stream-mode context refuses it."""
    import torch
    P = _particles()
    ctxs = [P.ParticleContext(sort_particles=True, initial_slot_capacity=1 << 16) for _ in range(2)]
    try:
        still = dict(drag=0.0, gravity=(0, 0, 0), lifeTime=1e9, spawnTime=1e9, timeStep=1 / 60)
        churn = dict(drag=0.1, gravity=(0, -9.81, 0), timeStep=1 / 60, emitPosition=dict(boxExtent=(3, 1, 3)),
                     emitVelocity=dict(sphere=dict(minRadius=0.5, maxRadius=2.0)))
        comps = [S.comp_c2(k, burst=700 + 100 * k, churn=True) for k in range(2)]
        comps += [abi.make_component(burstAmount=4000, spawnTime=0.002, lifeTime=0.05, lifeTimeVariance=0.1, **churn),   # one CTA, deaths and spawns
                  abi.make_component(burstAmount=60, **still),                                                      # one granule
                  abi.make_component(burstAmount=30000, emitPosition=dict(boxExtent=(4, 2, 4)), **still),            # tile path, 3 passes
                  abi.make_component(burstAmount=25000, emitPosition=dict(boxExtent=(0.002, 0.002, 0.002)), **still),  # tile path, 2 passes
                  abi.make_component(burstAmount=40000, spawnTime=1e-4, lifeTime=0.06, lifeTimeVariance=0.1, **churn)]  # tile path, deaths and spawns
        systems, idss = [], []
        for ctx in ctxs:
            s = ctx.create_system((1, 11, 3, 4))
            systems.append(s)
            idss.append([s.add_effect(c, abi.make_transform((k, 0, 300 if k == 5 else 0))) for k, c in enumerate(comps)])
        assert idss[0] == idss[1]
        ids = idss[0]
        ctxs[1].set_deferred_expansion(True)
        ctxs[1].set_assembly_share(1)   # ... and assembles through the bulk-asynchronous kernel, the other context through the load/store one
        cap = 1 << 17
        out = [torch.zeros((cap * 4, 8), dtype=torch.float32, device="cuda") for _ in range(2)]
        tot = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(2)]
        bufs = [ctx.run_buffer_create(cap) for ctx in ctxs]
        rec = torch.zeros((cap, 8), dtype=torch.float32, device="cuda")
        counts = []
        for f in range(8):
            fa = abi.make_frame_args(0.0175, f * 0.0175, f, camera=CAMERA)
            for ctx, s, o, t, b in zip(ctxs, systems, out, tot, bufs):
                s.update(ids, fa)
                ctx.run_buffer_pack(ctx.make_pack_list([s] * len(ids), ids), b)
                ctx.expand_from_peers([ctx.run_buffer_ptr(b)], o.data_ptr(), cap, t.data_ptr())
                ctx.synchronize()
            torch.cuda.synchronize()
            n = int(tot[0].item())
            assert n == int(tot[1].item()) and n > 50000
            assert torch.equal(out[0][:4 * n], out[1][:4 * n]), "frame %d: deferred assembly differs" % f
            # the unfused pair gives the same records
            total = ctxs[1].pack_runs([systems[1]] * len(ids), ids, rec.data_ptr(), cap)
            ctxs[1].synchronize(); torch.cuda.synchronize()
            assert total == n and torch.equal(rec[:n], out[0][:4 * n:4])
            for e in ids:   # the order read back per effect is the same in both contexts
                assert np.array_equal(systems[0].read_stream_slots(e), systems[1].read_stream_slots(e))
            counts.append(n)
        assert min(counts) < max(counts) - 1000, "the population should change over the frames (deaths in the tile-path effect)"
        ctxs[1].set_deferred_expansion(False)
        fa = abi.make_frame_args(0.0175, 8 * 0.0175, 8, camera=CAMERA)
        for s in systems:
            s.update(ids, fa)
        for e in ids:
            assert systems[0].read_stream(e).tobytes() == systems[1].read_stream(e).tobytes()
        plain = P.ParticleContext(initial_slot_capacity=1 << 12)
        try:
            with pytest.raises(Exception):
                plain.set_deferred_expansion(True)
        finally:
            plain.close()
    finally:
        for ctx in ctxs:
            ctx.close()


@pytest.mark.parametrize("attempt", range(4))
def test_type_and_effect_id_recycling(oracle_port, attempt):
    """Dense u32 ids recycled LIFO, types keyed by descriptor identity and refcounted
    (ParticleSystem.cpp:242-246, :682-728, :733-737, :796-806). Several attempts: a type / effect index
    that is freed and reused before the next submit is uploaded once, with its last value (two uploads
    of one index in the same parallel scatter were a race that only showed up now and then)."""
    P = _particles()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 16)
    try:
        s, o = ctx.create_system((1, 9, 3, 4)), orc.OracleSystem(oracle_port, (1, 9, 3, 4))
        comps = [S.comp_fountain(), S.comp_c5(), S.comp_point(), S.comp_timed_emitter(), S.comp_c1(burst=30)]
        log_s, log_o = [], []
        for sysm, log in ((s, log_s), (o, log_o)):
            log.append(sysm.reserve_type(comps[0]))          # 0
            log.append(sysm.reserve_type(comps[1]))          # 1
            log.append(sysm.reserve_type(comps[0]))          # 0 again (refcount 2)
            sysm.release_type(comps[0])
            sysm.release_type(comps[0])                      # type 0 freed
            log.append(sysm.reserve_type(comps[2]))          # reuses 0
            e = [sysm.add_effect(comps[k], abi.make_transform((k, 0, 0))) for k in (1, 2, 3)]
            log += e
            sysm.remove(e[1])
            sysm.remove(e[0])
            log.append(sysm.add_effect(comps[4], abi.make_transform()))   # reuses e[0]
            log.append(sysm.add_effect(comps[3], abi.make_transform()))   # reuses e[1]
            log.append(sysm.add_effect(comps[0], abi.make_transform()))   # new id
            ids = sorted(set(log[-3:] + [e[2]]))
            for f in range(10):
                sysm.update(ids, abi.make_frame_args(0.02, f * 0.02, f))
            log.append([sysm.effect_info(i).typeId for i in ids])
        assert log_s == log_o
        ids = sorted(set(log_s[-4:-1] + [log_s[6]]))
        S.assert_snapshots_equal(S.snapshot(o, ids), S.snapshot(s, ids), "after recycling")
    finally:
        ctx.close()


def test_update_with_nothing_active():
    P = _particles()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 16)
    try:
        s = ctx.create_system((1, 2, 3, 4))
        s.update([], abi.make_frame_args(0.016))
        e = s.add_effect(S.comp_point())
        s.update([], abi.make_frame_args(0.016))
        info = s.effect_info(e)
        assert (info.slotCount, info.aliveCount, info.simSteps, info.firstEmit) == (0, 0, 0, 1)
        recs, n = s.render([e], __import__("tests_render_args").make_render_args())
        assert recs == [] and n == 0
    finally:
        ctx.close()


@pytest.mark.parametrize("sort", [False, True], ids=["stream", "sorted"])
def test_c2_full_size_bit_exact(oracle_port, sort):
    """BASELINE configs[1] at its full size: 64 emitters x 16k particles, 4 effect types, one shared RNG,
    gravity/drag, bit-exact against the oracle (stream mode) / its stable depth sort (sort mode)."""
    P = _particles()
    ctx = P.ParticleContext(sort_particles=sort, initial_slot_capacity=1 << 21)
    try:
        seed = (1, 2, 3, 4)
        s, o = ctx.create_system(seed), orc.OracleSystem(oracle_port, seed)
        comps = [S.comp_c2(k, burst=16384, churn=True) for k in range(4)]
        ids = []
        for i in range(64):
            tr = abi.make_transform((4.0 * (i % 8), 0.0, 4.0 * (i // 8)))
            e1, e2 = s.add_effect(comps[i % 4], tr), o.add_effect(comps[i % 4], tr)
            assert e1 == e2
            ids.append(e1)
        for frame in range(5):
            fa = abi.make_frame_args(0.0175, game_time=frame * 0.0175, frame_index=frame, camera=CAMERA)
            s.update(ids, fa); o.update(ids, fa)
        assert s.rng_state() == o.rng_state()
        total = 0
        for e in ids:
            a, b = S.info_tuple(o.effect_info(e)), S.info_tuple(s.effect_info(e))
            assert a == b, "effect %d" % e
            total += a["aliveCount"]
            ref = o.read_stream(e)
            if sort:
                ref, _ = expected_sorted(ref)
            assert ref.tobytes() == s.read_stream(e).tobytes(), "effect %d stream" % e
            assert np.array_equal(o.read_free(e), s.read_free(e))
        assert total >= 64 * 16384
    finally:
        ctx.close()


# SPR_STRESS_SEEDS=n adds n more seeds (one-off soak runs on the GPU box; the default suite keeps three)
_RANDOM_SEEDS = [11, 12, 13] + list(range(100, 100 + int(os.environ.get("SPR_STRESS_SEEDS", "0"))))


@pytest.mark.parametrize("seed", _RANDOM_SEEDS)
@pytest.mark.parametrize("sort", [False, True], ids=["stream", "sorted"])
def test_random_api_sequence(oracle_port, seed, sort):
    """Randomised host-API sequence (add / remove / move / prepareForRemove / partial active lists / varying dt)
    mirrored on the oracle: exercises id recycling, slot-range reuse, effect relocation and pool growth."""
    import random
    rnd = random.Random(seed)
    P = _particles()
    ctx = P.ParticleContext(sort_particles=sort, initial_slot_capacity=1 << 12)
    try:
        sseed = (1, 100 + seed, 3, 4)
        s, o = ctx.create_system(sseed), orc.OracleSystem(oracle_port, sseed)
        comps = [S.comp_fountain(), S.comp_c5(), S.comp_point(), S.comp_timed_emitter(), S.comp_c1(burst=700), S.comp_local_space(),
                 S.comp_c2(1, burst=2500, churn=True), S.comp_c2(2, burst=90, churn=True), S.comp_no_emit()]
        live, transforms, t = [], {}, 0.0
        for frame in range(70):
            r = rnd.random()
            if (r < 0.25 and len(live) < 14) or not live:
                comp = rnd.choice(comps)
                tr = abi.make_transform((rnd.uniform(-5, 5), rnd.uniform(0, 2), rnd.uniform(-5, 5)), (0, 0, 0, 1), rnd.choice([1.0, 0.5, 2.0]))
                e1, e2 = s.add_effect(comp, tr), o.add_effect(comp, tr)
                assert e1 == e2
                live.append(e1); transforms[e1] = tr
            elif r < 0.35 and len(live) > 2:
                e = live.pop(rnd.randrange(len(live)))
                s.remove(e); o.remove(e)
                if r < 0.29:
                    # the freed id (and its slot range / table rows) is reused before the next submit
                    comp = rnd.choice(comps)
                    tr = abi.make_transform((rnd.uniform(-5, 5), rnd.uniform(0, 2), rnd.uniform(-5, 5)))
                    e1, e2 = s.add_effect(comp, tr), o.add_effect(comp, tr)
                    assert e1 == e2 == e
                    live.append(e1); transforms[e1] = tr
            elif r < 0.55:
                e = rnd.choice(live)
                tr = abi.make_transform((rnd.uniform(-5, 5), rnd.uniform(0, 2), rnd.uniform(-5, 5)), (0, 0.38268343, 0, 0.92387953), 1.0)
                u = S.transform_update(transforms[e], tr)
                s.update_transform(e, u); o.update_transform(e, u)
                transforms[e] = tr
            dt = rnd.choice([0.016, 0.02, 0.033, 0.005, 0.12])
            fa = abi.make_frame_args(dt, game_time=t, frame_index=frame, camera=CAMERA)
            if r > 0.9:
                e = rnd.choice(live)
                assert s.prepare_for_remove(e, fa) == o.prepare_for_remove(e, fa)
            active = [e for e in live if rnd.random() < 0.85]
            s.update(active, fa); o.update(active, fa)
            t += dt
            if frame % 7 == 6 or frame == 69:
                a, b = S.snapshot(o, live), S.snapshot(s, live)
                if sort:
                    assert a["rng"] == b["rng"]
                    for e in live:
                        assert a[e]["info"] == b[e]["info"], "frame %d effect %d" % (frame, e)
                        assert np.array_equal(a[e]["free"], b[e]["free"])
                        assert expected_sorted(a[e]["stream"])[0].tobytes() == b[e]["stream"].tobytes(), "frame %d effect %d" % (frame, e)
                else:
                    S.assert_snapshots_equal(a, b, "seed %d frame %d" % (seed, frame))
    finally:
        ctx.close()


def test_effects_from_json_descriptors(oracle_port):
    """Row N3 end to end: descriptors loaded from the reference's JSON prefab format drive the C ABI (the loaded
    descriptor's address is the effect type key) and simulate exactly like the oracle given the same values."""
    import json as _json
    P = _particles()
    text = [c for c in _json.load(open(os.path.join(os.path.dirname(__file__), "golden", "descriptors.json")))["cases"]
            if c["name"] == "canonical_compact"][0]["json"]
    ds = P.DescriptorSet(text)
    ctx = P.ParticleContext(initial_slot_capacity=1 << 12)
    try:
        assert len(ds) == 7
        seed = (1, 321, 3, 4)
        s, o = ctx.create_system(seed), orc.OracleSystem(oracle_port, seed)
        ids = []
        for i in range(len(ds)):
            tr = abi.make_transform((2.0 * i, 0.5, -1.0 * i))
            e1, e2 = s.add_effect(ds[i], tr), o.add_effect(ds[i], tr)
            assert e1 == e2
            ids.append(e1)
        assert s.add_effect(ds[0], abi.make_transform((9, 9, 9))) == o.add_effect(ds[0], abi.make_transform((9, 9, 9)))   # same descriptor -> same type
        ids.append(len(ids))
        assert s.effect_info(ids[-1]).typeId == s.effect_info(ids[0]).typeId
        for f in range(25):
            fa = abi.make_frame_args(0.02, game_time=f * 0.02, frame_index=f, camera=CAMERA)
            s.update(ids, fa); o.update(ids, fa)
        S.assert_snapshots_equal(S.snapshot(o, ids), S.snapshot(s, ids), "json descriptors")
    finally:
        ctx.close()
        ds.close()


def test_repeated_and_invalid_ids_are_refused(oracle_port):
    """ParticleSystem.cpp:812-821 would run a repeated id's sub-step loop once per entry; the device plan is per
    effect, so the call is refused (SPR_ERR_UNSUPPORTED = -5) BEFORE anything changes: the frames around the refused
    calls still match the oracle, which never sees them."""
    P = _particles()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 14)
    try:
        seed = (1, 2, 3, 4)
        s, o = ctx.create_system(seed), orc.OracleSystem(oracle_port, seed)
        s2 = ctx.create_system((5, 6, 7, 8))
        comp = S.comp_c1(burst=300)
        ids = [s.add_effect(comp, abi.make_transform((k, 0, 0))) for k in range(3)]
        assert ids == [o.add_effect(comp, abi.make_transform((k, 0, 0))) for k in range(3)]
        other = [s2.add_effect(comp, abi.make_transform((0, 0, 0)))]
        dt = s.effect_info(ids[0]).timeStep
        for frame in range(6):
            fa = abi.make_frame_args(dt, game_time=frame * dt, frame_index=frame)
            with pytest.raises(P.SpearError, match="status -5"):
                s.update([ids[0], ids[1], ids[0]], fa)
            with pytest.raises(P.SpearError, match="status -5"):
                ctx.update_batch([s, s2, s], [ids[:1], other, ids[1:]], fa)
            with pytest.raises(P.SpearError, match="status -1"):
                s.update([ids[0], 17], fa)
            with pytest.raises(P.SpearError, match="status -1"):
                ctx.update_batch([s, s2], [ids, [3]], fa)
            s.update(ids, fa); o.update(ids, fa)
            S.assert_snapshots_equal(S.snapshot(o, ids), S.snapshot(s, ids), "after refused calls, frame %d" % frame)
        # a removed id is refused as well, and accepted again once it is reused
        s.remove(ids[1]); o.remove(ids[1])
        fa = abi.make_frame_args(dt, game_time=6 * dt, frame_index=6)
        with pytest.raises(P.SpearError, match="status -1"):
            s.update(ids, fa)
        again = s.add_effect(comp, abi.make_transform((9, 0, 0)))
        assert again == o.add_effect(comp, abi.make_transform((9, 0, 0))) == ids[1]
        s.update(ids, fa); o.update(ids, fa)
        S.assert_snapshots_equal(S.snapshot(o, ids), S.snapshot(s, ids), "after id reuse")
    finally:
        ctx.close()


@pytest.mark.parametrize("sort", [False, True], ids=["stream", "sorted"])
def test_two_contexts_share_the_device(oracle_port, sort):
    """Two contexts on their own streams run their frames concurrently on one GPU. The chained scans hand their tiles
    out by ticket (the order the CTAs start), so neither context's look-back can wait for a CTA the other context's
    kernels keep off the SMs; both must match their oracles bit for bit."""
    P = _particles()
    ctxs = [P.ParticleContext(sort_particles=sort, initial_slot_capacity=1 << 16) for _ in range(2)]
    try:
        pairs = []
        for k, ctx in enumerate(ctxs):
            seed = (3, 100 + k, 5, 6)
            s, o = ctx.create_system(seed), orc.OracleSystem(oracle_port, seed)
            comp = comp_big(150_000 + 50_000 * k)
            assert s.add_effect(comp, abi.make_transform((1, 2, 3))) == o.add_effect(comp, abi.make_transform((1, 2, 3))) == 0
            pairs.append((s, o))
        dt = pairs[0][0].effect_info(0).timeStep
        for frame in range(8):
            fa = abi.make_frame_args(dt, game_time=frame * dt, frame_index=frame, camera=S.CAMERA)
            for s, _ in pairs:       # asynchronous: both contexts' kernels are in flight together
                s.update([0], fa)
            for _, o in pairs:
                o.update([0], fa)
        for k, (s, o) in enumerate(pairs):
            a, b = S.snapshot(o, [0]), S.snapshot(s, [0])
            assert a[0]["info"] == b[0]["info"] and a["rng"] == b["rng"]
            assert np.array_equal(a[0]["free"], b[0]["free"])
            exp = expected_sorted(a[0]["stream"])[0] if sort else a[0]["stream"]
            assert exp.tobytes() == b[0]["stream"].tobytes(), "context %d" % k
    finally:
        for ctx in ctxs:
            ctx.close()
