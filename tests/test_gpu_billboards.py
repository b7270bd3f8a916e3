"""GPU: sprBillboard* (spear_b200/csrc/billboards.cu) against the golden vectors of the
reference's cl::BillboardSystem and against the oracle on seeded frames: sort order, budget cut, crop, vertex
bytes and the atlas draw list, all bit-exact."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import billboards as B
from spear_b200 import abi
from tests_render_args import make_render_args

pytestmark = pytest.mark.gpu

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "billboards.json")))


def _particles():
    from spear_b200 import particles
    return particles


def test_billboards_match_golden_vectors():
    P = _particles()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 12)
    bs = P.BillboardSystem(ctx)
    ra = make_render_args()
    try:
        for frame in GOLDEN["frames"]:
            bs.add(B.random_frame(frame["seed"], frame["n"]))
            verts, draws, w2c = bs.render(ra)
            assert len(verts) // 4 == frame["quads"], "n=%d" % frame["n"]
            assert [list(d) for d in draws] == frame["draws"], "n=%d" % frame["n"]
            assert hashlib.sha256(verts.tobytes()).hexdigest() == frame["vertices_sha256"], "n=%d" % frame["n"]
            assert w2c == list(ra.worldToClip.v)
        verts, draws, _ = bs.render(ra)          # the frame was consumed (:227-228)
        assert len(verts) == 0 and draws == []
    finally:
        bs.close(); ctx.close()


@pytest.mark.parametrize("budget", [0, 100, 1 << 20])
def test_billboards_match_oracle_with_budgets(budget):
    """The frame budget is MaxBillboardsPerFrame = 4096 in the reference (:10); the port takes it as a parameter."""
    P = _particles()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 12)
    bs = P.BillboardSystem(ctx, max_billboards_per_frame=budget)
    ra = make_render_args()
    try:
        for seed, n in [(21, 1), (22, 33), (23, 1023), (24, 3000), (25, 70000)]:
            frame = B.random_frame(seed, n)
            for lo in range(0, n, 7000):          # several addBillboard batches per frame
                bs.add(frame[lo:lo + 7000])
            verts, draws, _ = bs.render(ra)
            pv, pd = B.port_render(frame, max_per_frame=budget or 4096)
            assert len(verts) == len(pv), "seed %d" % seed
            assert np.array_equal(verts.view(np.uint8), pv.view(np.uint8)), "seed %d" % seed
            assert draws == pd
    finally:
        bs.close(); ctx.close()


def test_billboards_all_equal_depth_keep_insertion_order():
    P = _particles()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 12)
    bs = P.BillboardSystem(ctx, max_billboards_per_frame=1 << 16)
    try:
        frame = [abi.make_billboard(atlas=1 + (i % 5), depth=3.5, transform=(1, 0, 0, 0, 1, 0, 0, 0, 1, float(i), 0, 0)) for i in range(5000)]
        bs.add(frame)
        verts, draws, _ = bs.render(make_render_args())
        assert len(verts) == 4 * 5000
        assert np.array_equal(verts["position"][::4, 0], np.arange(5000, dtype=np.float32) - 0.5)
        assert [d[0] for d in draws] == [1 + (i % 5) for i in range(5000)]
    finally:
        bs.close(); ctx.close()


def test_large_add_call_matches_chunked_adds():
    """sprBillboardAdd converts a call of 65536 billboards or more on several host threads (counts of loaded sprites
    per range, then the ranges side by side): the frame must be the one that the same billboards give when they are
    added in small calls, null sprites included."""
    P = _particles()
    base = B.random_frame(21, 4096)
    frame = base * 20                      # 81920 billboards, ~3 % of them without a sprite
    ra = make_render_args()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 12)
    one, many = P.BillboardSystem(ctx, max_billboards_per_frame=1 << 20), P.BillboardSystem(ctx, max_billboards_per_frame=1 << 20)
    try:
        one.add(frame)
        for k in range(0, len(frame), 20480):
            many.add(frame[k:k + 20480])
        v1, d1, _ = one.render(ra)
        v2, d2, _ = many.render(ra)
        assert len(v1) == len(v2) and len(v1) // 4 > 60000
        assert d1 == d2
        assert v1.tobytes() == v2.tobytes()
    finally:
        one.close(); many.close(); ctx.close()


def test_packed_billboards_from_host_and_device_memory_match_the_immediate_path():
    """sprBillboardAddPacked: the same frame through sprBillboardAdd, as one packed array in device memory (read in place),
    as a packed host array, and mixed with sprBillboardAdd calls in between (the frame is the calls in call order) gives
    the same vertex bytes and draw list -- which the golden test ties to the reference."""
    import ctypes as C
    import torch
    P = _particles()
    ctx = P.ParticleContext(initial_slot_capacity=1 << 12)
    ra = make_render_args()
    try:
        for budget, (seed, n) in [(0, (31, 5000)), (1 << 20, (32, 70000)), (100, (33, 777))]:
            bs = P.BillboardSystem(ctx, max_billboards_per_frame=budget)
            try:
                frame = B.random_frame(seed, n)
                bs.add(frame)
                want_v, want_d, _ = bs.render(ra)
                packed, m = P.BillboardSystem.pack(frame)
                assert C.sizeof(abi.BillboardPacked) == 120 and 0 < m <= n
                host = np.frombuffer(packed, dtype=np.uint8, count=m * 120).copy()
                dev = torch.from_numpy(host).cuda()
                torch.cuda.synchronize()
                for variant in ("device", "host", "mixed"):
                    if variant == "device":
                        bs.add_packed(dev.data_ptr(), m, device=True)
                    elif variant == "host":
                        bs.add_packed(host.ctypes.data, m, device=False)
                    else:
                        # first third through the immediate path, the middle from device memory, the rest from host memory
                        a = n // 3
                        pa = P.BillboardSystem.pack(frame[:a])[1]            # packed entries the first third makes
                        pb = pa + (m - pa) // 2
                        bs.add(frame[:a])
                        bs.add_packed(dev.data_ptr() + pa * 120, pb - pa, device=True)
                        bs.add_packed(host.ctypes.data + pb * 120, m - pb, device=False)
                    v, d, _ = bs.render(ra)
                    assert d == want_d, "%s n=%d" % (variant, n)
                    assert np.array_equal(v.view(np.uint8), want_v.view(np.uint8)), "%s n=%d" % (variant, n)
                v, d, _ = bs.render(ra)
                assert len(v) == 0 and d == []
            finally:
                bs.close()
    finally:
        ctx.close()
