#!/usr/bin/env python3
"""Full-size golden digests from the REFERENCE ITSELF (oracle/_ref/libspear_ref.so), for the sizes BASELINE.json
names and bench.py measures. Run in the build container only (about five minutes of CPU):

    python tests/golden/make_golden_full.py

  c3   BASELINE configs[2]: one effect, 10M particles (bench.py's C3 descriptor), 3 frames
  c4   BASELINE configs[3], one GPU's slice: 128 independent systems x 65536 particles, 3 frames
  c5   BASELINE configs[4], one GPU's slice: 1000 systems x 100 high-churn effects, 20 frames

Per frame the file holds sha256 digests of what the reference computed: the vertex stream in the reference's own
order (ascending slot) and in per-particle depth order (stable sort of that stream, tests/scenarios.py expected_sorted),
the free lists, the counts and the RNG states. tests/test_gpu_full_size.py drives the CUDA path through the same
frames on the GPU box (where /root/reference does not exist) and compares digests.
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import scenarios as S  # noqa: E402
from spear_b200 import abi  # noqa: E402

CAMERA = S.CAMERA
REC = np.dtype([("pos", "<f4", 3), ("life", "<f4"), ("vel", "<f4", 3), ("seed", "<f4")])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------------------
# workload definitions shared by the generator and the GPU test (same descriptors as bench.py)
# ---------------------------------------------------------------------------

def c3_effect():
    import bench
    wl = bench.workload_components("c3")
    return wl["effects"][0]


def c4_systems(n=128):
    """[(seed, component, transform)] of one GPU's C4 slice (bench.py build_ours, rank 0)."""
    import bench
    wl = bench.workload_components("c4")
    return [((1, 2 + i, 3, 4), wl["comps"][i % len(wl["comps"])], abi.make_transform((4.0 * (i % 16), 0.0, 4.0 * (i // 16)))) for i in range(n)]


def c5_systems(n=1000, per_system=100):
    """[(seed, component, [transforms])] of one GPU's C5 slice."""
    comp = S.comp_c5()
    return [((1, 2 + i, 3, 4), comp, [abi.make_transform((4.0 * (i % 16) + 0.25 * k, 0.0, 4.0 * (i // 16))) for k in range(per_system)]) for i in range(n)]


def sorted_records(records, segment=None):
    """Depth order of un-expanded stream records: back to front, ties by stream position; per segment when given."""
    key = S.sort_key(records, CAMERA)
    order = np.argsort(key, kind="stable") if segment is None else np.lexsort((key, segment))
    return records[order]


# ---------------------------------------------------------------------------

def golden_c3(orc):
    comp, tr = c3_effect()
    o = orc.OracleSystem("reference", (1, 2, 3, 4))
    e = o.add_effect(comp, tr)
    dt = o.effect_info(e).timeStep
    frames = []
    for f in range(3):
        o.update([e], abi.make_frame_args(dt, game_time=f * dt, frame_index=f, camera=CAMERA))
        info = o.effect_info(e)
        stream = o.read_stream(e)
        recs = stream[::4]
        assert np.array_equal(stream[1::4], recs) and np.array_equal(stream[3::4], recs)
        frames.append(dict(slotCount=info.slotCount, aliveCount=info.aliveCount, freeCount=info.freeCount, rng="%016x" % o.rng_state()[0],
                           free=sha(o.read_free(e)), stream=sha(stream), sorted=sha(np.repeat(sorted_records(recs), 4))))
        print("c3 frame", f, frames[-1]["aliveCount"], flush=True)
    o.close()
    return dict(dtBits=S.f32_bits(dt), frames=frames)


def golden_c4(orc):
    systems = []
    for seed, comp, tr in c4_systems():
        o = orc.OracleSystem("reference", seed)
        systems.append((o, o.add_effect(comp, tr)))
    dt = max(o.effect_info(e).timeStep for o, e in systems)
    frames = []
    for f in range(3):
        fa = abi.make_frame_args(dt, game_time=f * dt, frame_index=f, camera=CAMERA)
        recs, frees, counts, rngs = [], [], [], []
        for o, e in systems:
            o.update([e], fa)
            info = o.effect_info(e)
            counts.append((info.slotCount, info.aliveCount, info.freeCount))
            rngs.append(o.rng_state()[0])
            frees.append(o.read_free(e))
            recs.append(sorted_records(o.read_stream(e)[::4]))
        frames.append(dict(total=int(sum(c[1] for c in counts)), counts=sha(np.array(counts, dtype=np.uint32)), rng=sha(np.array(rngs, dtype=np.uint64)),
                           free=sha(np.concatenate(frees)), sorted=sha(np.concatenate(recs))))
        print("c4 frame", f, frames[-1]["total"], flush=True)
    for o, _ in systems:
        o.close()
    return dict(dtBits=S.f32_bits(dt), frames=frames)


C5_SAMPLE_STRIDE = 997


def golden_c5(orc):
    systems = []
    for seed, comp, trs in c5_systems():
        o = orc.OracleSystem("reference", seed)
        systems.append((o, [o.add_effect(comp, tr) for tr in trs]))
    dt = max(o.effect_info(ids[0]).timeStep for o, ids in systems) * 1.1
    frames = []
    for f in range(20):
        fa = abi.make_frame_args(dt, game_time=f * dt, frame_index=f, camera=CAMERA)
        recs, seg, alive, rngs, samples = [], [], [], [], {}
        k = 0
        for o, ids in systems:
            o.update(ids, fa)
            rngs.append(o.rng_state()[0])
            for e in ids:
                r = o.read_stream(e)[::4]
                recs.append(r)
                seg.append(np.full(len(r), k, dtype=np.uint32))
                alive.append(len(r))
                if k % C5_SAMPLE_STRIDE == 0:
                    info = o.effect_info(e)
                    samples[str(k)] = dict(slotCount=info.slotCount, aliveCount=info.aliveCount, freeCount=info.freeCount,
                                           spawnTimer=S.f32_bits(info.spawnTimer), free=sha(o.read_free(e)))
                k += 1
        recs = np.concatenate(recs)
        frames.append(dict(total=int(len(recs)), alive=sha(np.array(alive, dtype=np.uint32)), rng=sha(np.array(rngs, dtype=np.uint64)),
                           stream=sha(recs), sorted=sha(sorted_records(recs, np.concatenate(seg))), samples=samples))
        print("c5 frame", f, frames[-1]["total"], flush=True)
    for o, _ in systems:
        o.close()
    return dict(dtBits=S.f32_bits(dt), frames=frames)


def main():
    from oracle import build_ref, orc
    build_ref.build(verbose=False)
    out = {}
    for name, fn in (("c3", golden_c3), ("c4", golden_c4), ("c5", golden_c5)):
        t0 = time.time()
        out[name] = fn(orc)
        print("%s: %.0f s" % (name, time.time() - t0), flush=True)
    with open(os.path.join(HERE, "full_size.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote full_size.json")


if __name__ == "__main__":
    main()
