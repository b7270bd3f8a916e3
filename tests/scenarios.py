"""Synthetic effect descriptors and a scenario driver shared by the oracle tests,
the GPU parity tests and the golden-vector generator.

Every descriptor is synthetic (no private assets); unspecified fields keep the
reference defaults (src/server/ServerState.h:187-239). Scenario names follow
BASELINE.json's configs (C1..C5) plus edge cases the hot path has:
bursts, the 20-spawn cap, emitterOnTime, localSpace, rotation, attractor,
gravity points, moving emitters, multi-substep frames, effect removal/reuse.
"""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spear_b200 import abi  # noqa: E402


def comp_c1(burst=10000):
    """BASELINE.json configs[0] / SURVEY.md section 8(d) C1."""
    return abi.make_component(
        timeStep=1 / 60, burstAmount=burst, spawnTime=0.001, spawnTimeVariance=0.0005,
        lifeTime=4.0, lifeTimeVariance=2.0, drag=0.3, gravity=(0, -9.81, 0),
        emitPosition=dict(boxExtent=(1, 0.5, 1)),
        emitVelocity=dict(sphere=dict(minRadius=1, maxRadius=3)),
        scaleSpline=[(0, 0), (0.3, 1), (1, 0.2)],
        gradient=dict(points=[(0, (1, 0.5, 0.1)), (1, (0.1, 0.1, 0.1))]))


def comp_c2(kind, burst=16384, churn=False):
    """4 effect types cycling (box/sphere emitters, with/without rotation)."""
    kw = dict(
        timeStep=1 / 60, burstAmount=burst, spawnTime=1e9,
        lifeTime=8.0 if churn else 1e9, lifeTimeVariance=4.0 if churn else 0.0,
        drag=0.2, gravity=(0, -2.0, 0), renderOrder=float(kind & 1),
        scaleSpline=[(0, 0.2), (0.5, 1), (1, 0)], alphaSpline=[(0, 0), (0.1, 1), (1, 0)],
        additiveSpline=[(0, 1), (1, 0)], erosionSpline=[(0, 0), (1, 1)],
        gradient=dict(points=[(0, (1, 0.9, 0.4)), (0.5, (1, 0.3, 0.1)), (1, (0.05, 0.05, 0.05))]))
    if kind == 0:
        kw.update(emitPosition=dict(boxExtent=(2, 0.2, 2)), emitVelocity=dict(offset=(0, 2, 0), boxExtent=(1, 1, 1)))
    elif kind == 1:
        kw.update(emitPosition=dict(sphere=dict(minRadius=0.2, maxRadius=1.0)),
                  emitVelocity=dict(sphere=dict(minRadius=0.5, maxRadius=2.0, maxPhi=90.0)))
    elif kind == 2:
        kw.update(emitPosition=dict(boxExtent=(1, 1, 1), rotation=(0, 45, 30)),
                  emitVelocity=dict(offset=(0, 1, 0), sphere=dict(minRadius=0.1, maxRadius=1.5)))
    else:
        kw.update(emitPosition=dict(sphere=dict(minRadius=0.0, maxRadius=1.5, scale=(1, 0.3, 1)), rotation=(20, 0, 0)),
                  emitVelocity=dict(boxExtent=(3, 0.5, 3), rotation=(0, 90, 0)))
    return abi.make_component(**kw)


def comp_c5():
    """20 die + 20 spawn per step, ~40 alive, 60 slots."""
    return abi.make_component(
        timeStep=1 / 60, spawnTime=0.0, lifeTime=2.0 / 60.0, lifeTimeVariance=0.0,
        gravity=(0, -9.81, 0), drag=0.1,
        emitPosition=dict(boxExtent=(1, 1, 1)), emitVelocity=dict(sphere=dict(minRadius=1, maxRadius=2)))


def comp_fountain():
    """Timer emission with variance, sphere position + rotation + attractor (SURVEY.md Appendix C probe 2)."""
    return abi.make_component(
        timeStep=1 / 30, burstAmount=37, spawnTime=0.004, spawnTimeVariance=0.003,
        lifeTime=0.8, lifeTimeVariance=0.6, drag=0.5, gravity=(0.5, -9.81, 0.25),
        emitPosition=dict(offset=(0, 0.5, 0), sphere=dict(minRadius=0.2, maxRadius=0.8, minTheta=30, maxTheta=300, minPhi=10, maxPhi=120, scale=(1, 0.5, 2)), rotation=(10, 20, 30)),
        emitVelocity=dict(offset=(0, 3, 0), boxExtent=(0.5, 0.5, 0.5), sphere=dict(minRadius=0.5, maxRadius=1.5)),
        emitVelocityAttractorOffset=(0, 1, 0), emitVelocityAttractorStrength=-1.5,
        cullPadding=0.75)


def comp_gravity_points():
    return abi.make_component(
        timeStep=1 / 60, burstAmount=500, spawnTime=0.01, lifeTime=2.0, lifeTimeVariance=1.0, drag=0.1,
        gravity=(0, -1, 0),
        emitPosition=dict(boxExtent=(2, 2, 2)), emitVelocity=dict(sphere=dict(minRadius=0.5, maxRadius=1.0)),
        gravityPoints=[dict(position=(0, 1, 0), radius=0.5, strength=3.0), dict(position=(1, 0, 1), radius=0.01, strength=-2.0)])


def comp_local_space():
    return abi.make_component(
        timeStep=1 / 60, burstAmount=64, spawnTime=0.02, lifeTime=1.0, localSpace=1, cullPadding=0.5,
        emitPosition=dict(boxExtent=(1, 1, 1)), emitVelocity=dict(offset=(0, 1, 0), boxExtent=(1, 0, 1)))


def comp_timed_emitter():
    """emitterOnTime expiry (ParticleSystem.cpp:466-471) and short lives so the effect drains."""
    return abi.make_component(
        timeStep=1 / 60, burstAmount=5, spawnTime=0.005, emitterOnTime=0.1, lifeTime=0.15, lifeTimeVariance=0.1,
        emitPosition=dict(boxExtent=(1, 1, 1)), emitVelocity=dict(boxExtent=(2, 2, 2)))


def comp_no_emit():
    """emitterOnTime = 0 -> only the burst is emitted."""
    return abi.make_component(timeStep=0.05, burstAmount=9, spawnTime=0.01, emitterOnTime=0.0, lifeTime=0.3,
                              emitPosition=dict(boxExtent=(1, 1, 1)))


def comp_point():
    """No box, no sphere: K = 2 draws per timer spawn."""
    return abi.make_component(timeStep=0.02, spawnTime=0.0015, lifeTime=0.1, emitVelocity=dict(offset=(0, 1, 0)))


def f32_bits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


def info_tuple(info):
    """Bit-exact comparable view of SprEffectInfo."""
    b = info.gpuBounds
    return dict(
        typeId=info.typeId, slotCount=info.slotCount, aliveCount=info.aliveCount, freeCount=info.freeCount,
        timeStep=f32_bits(info.timeStep), timeDelta=f32_bits(info.timeDelta), spawnTimer=f32_bits(info.spawnTimer),
        emitOnTimer=f32_bits(info.emitOnTimer), stopEmit=info.stopEmit, firstEmit=info.firstEmit,
        bounds=tuple(f32_bits(v) for v in (b.origin.x, b.origin.y, b.origin.z, b.extent.x, b.extent.y, b.extent.z)),
        simSteps=info.simSteps)


def snapshot(system, effect_ids):
    """Everything SURVEY.md section 8(c) lists, per effect, as bytes/ints."""
    snap = {"rng": system.rng_state()}
    for e in effect_ids:
        info = info_tuple(system.effect_info(e))
        if info["aliveCount"] == 0:
            info["bounds"] = None  # +-inf arithmetic on an empty effect; never read by the reference (:839)
        snap[e] = dict(info=info, free=system.read_free(e).copy(), stream=system.read_stream(e).copy())
    return snap


def assert_snapshots_equal(a, b, where="", float_exact=True, rtol=0.0):
    assert a["rng"] == b["rng"], "%s rng state %r != %r" % (where, a["rng"], b["rng"])
    for e in a:
        if e == "rng":
            continue
        ia, ib = dict(a[e]["info"]), dict(b[e]["info"])
        if not float_exact:
            ia.pop("bounds"); ib.pop("bounds")
        assert ia == ib, "%s effect %d info\n%r\n%r" % (where, e, ia, ib)
        assert np.array_equal(a[e]["free"], b[e]["free"]), "%s effect %d freeIndices differ" % (where, e)
        sa, sb = a[e]["stream"], b[e]["stream"]
        assert len(sa) == len(sb), "%s effect %d stream length %d != %d" % (where, e, len(sa), len(sb))
        if float_exact:
            assert sa.tobytes() == sb.tobytes(), "%s effect %d vertex stream bytes differ" % (where, e)
        else:
            fa = sa.view("<f4").reshape(len(sa), 8)
            fb = sb.view("<f4").reshape(len(sb), 8)
            assert np.array_equal(fa[:, 7], fb[:, 7]), "%s effect %d seeds differ" % (where, e)
            np.testing.assert_allclose(fa, fb, rtol=rtol, atol=rtol)


class Scenario:
    """A deterministic script: effects to add and per-frame actions."""

    def __init__(self, name, seed, effects, frames, dt=None, moves=None, removes=None, prepare=None, readds=None, float_exact=True):
        self.name = name
        self.seed = seed
        self.effects = effects      # list of (component, transform)
        self.frames = frames
        self.dt = dt                # None -> timeStep of effect 0 (exactly one sim step per frame for it)
        self.moves = moves or {}    # frame -> list of (effect index, SprTransform)
        self.removes = removes or {}   # frame -> list of effect ids to remove
        self.prepare = prepare or {}   # frame -> list of effect ids to prepareForRemove
        self.readds = readds or {}     # frame -> list of (component, transform) to add
        self.float_exact = float_exact


def transform_update(prev, cur):
    """cl::TransformUpdate for a transform change; entityToWorld = Transform::asMatrix() (Transform.h:15-17)."""
    u = abi.TransformUpdate()
    u.previousTransform = prev
    u.transform = cur
    q, s2 = cur.rotation, 2.0 * cur.scale
    f = np.float32
    x, y, z, w = f(q.x), f(q.y), f(q.z), f(q.w)
    s2 = f(2.0) * f(cur.scale)
    xx, xy, xz, xw = x * x, x * y, x * z, x * w
    yy, yz, yw = y * y, y * z, y * w
    zz, zw = z * z, z * w
    h = f(0.5)
    m = [s2 * (-yy - zz + h), s2 * (xy + zw), s2 * (-yw + xz),
         s2 * (-zw + xy), s2 * (-xx - zz + h), s2 * (xw + yz),
         s2 * (xz + yw), s2 * (-xw + yz), s2 * (-xx - yy + h),
         cur.position.x, cur.position.y, cur.position.z]
    for i in range(12):
        u.entityToWorld.v[i] = float(m[i])
    return u


def run_scenario(system, sc, on_frame=None):
    """Drive `system` (OracleSystem or the product's ParticleSystem wrapper) through `sc`.
    Returns the list of per-frame snapshots."""
    ids = []
    transforms = {}
    for comp, tr in sc.effects:
        e = system.add_effect(comp, tr)
        ids.append(e)
        transforms[e] = tr
    dt = sc.dt if sc.dt is not None else system.effect_info(ids[0]).timeStep
    snaps = []
    game_time = 0.0
    for frame in range(sc.frames):
        for (e, tr) in sc.moves.get(frame, []):
            system.update_transform(e, transform_update(transforms[e], tr))
            transforms[e] = tr
        for (comp, tr) in sc.readds.get(frame, []):
            e = system.add_effect(comp, tr)
            if e not in ids:
                ids.append(e)
            transforms[e] = tr
        fa = abi.make_frame_args(dt, game_time=game_time, frame_index=frame, camera=(16.0, 20.0, -30.0))
        for e in sc.prepare.get(frame, []):
            system.prepare_for_remove(e, fa)
        for e in sc.removes.get(frame, []):
            system.remove(e)
            ids.remove(e)
        system.update(ids, fa)
        game_time += dt
        snap = snapshot(system, ids)
        snaps.append(snap)
        if on_frame is not None:
            on_frame(frame, snap)
    return snaps


def _tr(p=(0, 0, 0), r=(0, 0, 0, 1), s=1.0):
    return abi.make_transform(position=p, rotation=r, scale=s)


def scenarios():
    """Small scenarios the CPU oracles finish in well under a second each."""
    q = (0.18257418, 0.36514837, 0.54772256, 0.73029674)  # unit quaternion
    out = []
    out.append(Scenario("c1_small", (1, 2, 3, 4), [(comp_c1(burst=1000), _tr((1, 2, 3)))], frames=90))
    out.append(Scenario("burst6", (1, 2, 3, 4), [(abi.make_component(
        timeStep=0.1, burstAmount=6, spawnTime=1000.0, emitPosition=dict(boxExtent=(1, 1, 1)),
        emitVelocity=dict(sphere=dict(minRadius=1, maxRadius=2))), _tr())], frames=3))
    out.append(Scenario("fountain", (9, 77, 3, 4), [(comp_fountain(), _tr((1, 0, -2), q, 1.5))], frames=80))
    out.append(Scenario("timed", (5, 6, 7, 8), [(comp_timed_emitter(), _tr((0, 1, 0)))], frames=40))
    out.append(Scenario("noemit", (5, 6, 7, 8), [(comp_no_emit(), _tr())], frames=12))
    out.append(Scenario("point", (5, 123456789, 7, 8), [(comp_point(), _tr((3, 3, 3), q, 0.5))], frames=30))
    out.append(Scenario("local_space", (1, 42, 3, 4), [(comp_local_space(), _tr((5, 1, 2), q, 2.0))], frames=40,
                        moves={10: [(0, _tr((6, 1, 2), q, 2.0))], 11: [(0, _tr((7, 2, 2), (0, 0, 0, 1), 1.0))]}))
    out.append(Scenario("moving", (1, 43, 3, 4), [(comp_fountain(), _tr((0, 0, 0)))], frames=30,
                        moves={f: [(0, _tr((0.3 * f, 0.1 * f, 0), q if f % 2 else (0, 0, 0, 1), 1.0 + 0.01 * f))] for f in range(1, 30)}))
    out.append(Scenario("gravity_points", (1, 44, 3, 4), [(comp_gravity_points(), _tr((0.5, 0, 0), q, 1.0))], frames=40,
                        float_exact=False))
    # several effects sharing one RNG, different types, a frame dt that gives 0/1/2 substeps per effect
    multi = [(comp_c2(k % 4, burst=300 + 17 * k, churn=True), _tr((4 * (k % 4), 0, 4 * (k // 4)))) for k in range(8)]
    multi.append((comp_fountain(), _tr((1, 1, 1))))
    multi.append((comp_c5(), _tr((2, 2, 2))))
    out.append(Scenario("multi", (1, 2024, 3, 4), multi, frames=50, dt=0.0171))
    out.append(Scenario("multi_bigdt", (1, 2025, 3, 4), multi[:5] + [(comp_c5(), _tr())], frames=12, dt=0.25))
    out.append(Scenario("churn_c5", (1, 7, 3, 4), [(comp_c5(), _tr((k, 0, 0))) for k in range(12)], frames=20))
    # nothing to simulate: no burst and a first timer spawn far in the future, then a dt of 0
    idle = abi.make_component(timeStep=0.02, spawnTime=5.0, lifeTime=0.5, emitPosition=dict(boxExtent=(1, 1, 1)))
    out.append(Scenario("idle_then_zero_dt", (1, 3, 3, 4), [(idle, _tr()), (comp_point(), _tr((1, 1, 1)))], frames=8, dt=0.0))
    # the frame dt is clamped to 0.1 (ParticleSystem.cpp:810): many sub-steps per frame with a small timeStep
    fast = abi.make_component(timeStep=0.004, burstAmount=40, spawnTime=0.0007, spawnTimeVariance=0.0004, lifeTime=0.05,
                              lifeTimeVariance=0.05, gravity=(0, -9.81, 0), emitPosition=dict(sphere=dict(minRadius=0.1, maxRadius=0.5)),
                              emitVelocity=dict(boxExtent=(2, 2, 2)))
    out.append(Scenario("clamped_dt_many_substeps", (1, 31, 3, 4), [(fast, _tr((0, 1, 0), q, 1.0)), (comp_c5(), _tr((1, 0, 0)))], frames=6, dt=1.0))
    # removal / id reuse / prepareForRemove
    out.append(Scenario("remove_reuse", (1, 99, 3, 4),
                        [(comp_fountain(), _tr()), (comp_c5(), _tr((1, 0, 0))), (comp_timed_emitter(), _tr((2, 0, 0)))],
                        frames=40, prepare={8: [0]}, removes={20: [1]},
                        readds={25: [(comp_c1(burst=50), _tr((3, 0, 0)))], 30: [(comp_c5(), _tr((4, 0, 0)))]}))
    return out


# ---------------------------------------------------------------------------
# Per-particle depth order (extension): oracle = stable sort of the
# reference's own compacted stream.
# ---------------------------------------------------------------------------

CAMERA = (16.0, 20.0, -30.0)


def sort_key(stream_records, camera=CAMERA):
    """~enc(|camera - pos|^2) with float32 arithmetic in the reference's order (src/sf/Vector.h:99)."""
    pos = stream_records["pos"].astype(np.float32)
    cam = np.asarray(camera, dtype=np.float32)
    d = cam[None, :] - pos
    d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
    bits = d2.view(np.uint32)
    enc = np.where(bits & 0x80000000, ~bits, bits | np.uint32(0x80000000)).astype(np.uint32)
    return ~enc


def expected_sorted(oracle_stream, camera=CAMERA):
    """Back-to-front, ties by ascending slot (BillboardOrder idiom, src/client/BillboardSystem.cpp:41-51)."""
    recs = oracle_stream[::4]
    order = np.argsort(sort_key(recs, camera), kind="stable")
    return np.repeat(recs[order], 4), order


# ---------------------------------------------------------------------------
# Digests for the golden vectors (tests/golden/)
# ---------------------------------------------------------------------------

def snapshot_digest(snap):
    import hashlib
    out = {"rng": "%016x" % snap["rng"][0]}
    for e in sorted(k for k in snap if k != "rng"):
        info = dict(snap[e]["info"])
        info["bounds"] = list(info["bounds"]) if info["bounds"] is not None else None
        out[str(e)] = dict(info=info,
                           free=hashlib.sha256(snap[e]["free"].tobytes()).hexdigest()[:16],
                           stream=hashlib.sha256(snap[e]["stream"].tobytes()).hexdigest()[:16])
    return out
