"""TEST INFRASTRUCTURE ONLY: the CPU oracles of SURVEY.md 8(f) row N4 (cl::BillboardSystem).

  BillboardOracle   ctypes loader of oracle/_ref/libspear_ref_billboard.so = the reference's own
                    src/client/BillboardSystem.cpp behind oracle/ref_billboard_harness.cpp
  port_render()     numpy float32 restatement of BillboardSystem::renderMain (:119-228) and of
                    addBillboard's packColor (:12-28), each step citing the lines it follows.
Parity status: PINNED -- tests/test_billboard_oracle.py checks the port against the reference build on
seeded random frames, and both against tests/golden/billboards.json (generated from the reference build
by tests/golden/make_golden_billboards.py; the reference tree does not exist on the GPU box).

Never imported by the product package (spear_b200); used by tests/ only.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from spear_b200 import abi  # noqa: E402

LIB = os.path.join(HERE, "_ref", "libspear_ref_billboard.so")
F = np.float32


def have_reference():
    return os.path.isfile(LIB)


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB)
        vp, u32 = C.c_void_p, C.c_uint32
        lib.orc_bb_create.restype = vp
        lib.orc_bb_destroy.argtypes = [vp]
        lib.orc_bb_add.argtypes = [vp, C.POINTER(abi.Billboard), u32]
        lib.orc_bb_render.restype = u32
        lib.orc_bb_render.argtypes = [vp, C.POINTER(abi.RenderArgs), vp, u32, C.POINTER(abi.BillboardDraw), u32, C.POINTER(u32), C.POINTER(C.c_float * 16)]
        _lib = lib
    return _lib


class BillboardOracle:
    """The reference's cl::BillboardSystem run on the CPU."""

    def __init__(self):
        self.lib = _load()
        self.h = self.lib.orc_bb_create()

    def close(self):
        if self.h:
            self.lib.orc_bb_destroy(self.h)
            self.h = None

    def add(self, billboards):
        n = len(billboards)
        arr = (abi.Billboard * max(n, 1))(*billboards)
        self.lib.orc_bb_add(self.h, arr, n)

    def render(self, render_args, capacity=1 << 16):
        verts = np.zeros(4 * capacity, dtype=abi.BILLBOARD_VERTEX_DTYPE)
        draws = (abi.BillboardDraw * 4096)()
        nd = C.c_uint32(0)
        w2c = (C.c_float * 16)()
        q = self.lib.orc_bb_render(self.h, C.byref(render_args), verts.ctypes.data, capacity, draws, 4096, C.byref(nd), C.byref(w2c))
        assert q <= capacity and nd.value <= 4096
        return verts[:4 * q].copy(), [(draws[i].atlas, draws[i].count, draws[i].firstQuad) for i in range(nd.value)], list(w2c)


def pack_channel(f):
    """packChannel, BillboardSystem.cpp:12-15: sf::clamp((uint32_t)(f * 255.0f), 0u, 255u); the conversion of an
    out-of-range float is what g++ emits on x86-64 (low 32 bits of the truncating 64-bit conversion)."""
    v = F(f) * F(255.0)
    if np.isnan(v) or not (-9223372036854775808.0 <= float(v) < 9223372036854775808.0):
        u = 0
    else:
        u = int(np.trunc(float(v))) & 0xFFFFFFFF
    return min(u, 255)


def pack_color(c):
    return pack_channel(c[0]) | pack_channel(c[1]) << 8 | pack_channel(c[2]) << 16 | pack_channel(c[3]) << 24


def port_render(billboards, max_per_frame=4096):
    """billboards: list of abi.Billboard in insertion order (null sprites, atlas == 0, are dropped like addBillboard
    does, :82). Returns (vertices structured array [4 * quads], draws [(atlas, count, firstQuad)])."""
    bbs = [b for b in billboards if b.sprite.atlas]
    depth = np.array([b.depth for b in bbs], dtype=F)
    order = np.argsort(depth + F(0.0), kind="stable") if len(bbs) else np.zeros(0, dtype=np.int64)   # (depth, index), :41-51, :119
    order = order[:max_per_frame]                                                                      # :121-123
    verts = []
    draws = []
    for i in order:
        b = bbs[int(i)]
        sp = b.sprite
        # sf::max / sf::min are _mm_max_ss / _mm_min_ss: the second operand on ties (src/sf/Base.h:107-108)
        mn = [F(sp.minVert.x) if F(sp.minVert.x) > F(b.cropMin.x) else F(b.cropMin.x), F(sp.minVert.y) if F(sp.minVert.y) > F(b.cropMin.y) else F(b.cropMin.y)]   # :136
        mx = [F(sp.maxVert.x) if F(sp.maxVert.x) < F(b.cropMax.x) else F(b.cropMax.x), F(sp.maxVert.y) if F(sp.maxVert.y) < F(b.cropMax.y) else F(b.cropMax.y)]   # :137
        if mn[0] >= mx[0] or mn[1] >= mx[1]:                                                           # :138
            continue
        uv_min = (mn[0] * F(sp.vertUvScale.x) + F(sp.vertUvBias.x), mn[1] * F(sp.vertUvScale.y) + F(sp.vertUvBias.y))   # :140-143
        uv_max = (mx[0] * F(sp.vertUvScale.x) + F(sp.vertUvBias.x), mx[1] * F(sp.vertUvScale.y) + F(sp.vertUvBias.y))
        mn = [mn[0] - F(b.anchor.x), mn[1] - F(b.anchor.y)]                                            # :145-146
        mx = [mx[0] - F(b.anchor.x), mx[1] - F(b.anchor.y)]
        t = [F(v) for v in b.transform]
        m00, m10, m20, m01, m11, m21, m03, m13, m23 = t[0], t[1], t[2], t[3], t[4], t[5], t[9], t[10], t[11]
        xx0, xx1 = mn[0] * m00 + m03, mx[0] * m00 + m03                                                # :148-161
        yy0, yy1 = mn[1] * m11 + m13, mx[1] * m11 + m13
        xy0, xy1 = mn[1] * m01, mx[1] * m01
        yx0, yx1 = mn[0] * m10, mx[0] * m10
        zx0, zx1 = mn[0] * m20 + m23, mx[0] * m20 + m23
        zy0, zy1 = mn[1] * m21, mx[1] * m21
        col = pack_color((b.color.x, b.color.y, b.color.z, b.color.w))
        verts += [((xx0 + xy0, yx0 + yy0, zx0 + zy0), (uv_min[0], uv_min[1]), col),                     # :165-191
                  ((xx1 + xy0, yx1 + yy0, zx1 + zy0), (uv_max[0], uv_min[1]), col),
                  ((xx0 + xy1, yx0 + yy1, zx0 + zy1), (uv_min[0], uv_max[1]), col),
                  ((xx1 + xy1, yx1 + yy1, zx1 + zy1), (uv_max[0], uv_max[1]), col)]
        if draws and draws[-1][0] == sp.atlas:                                                         # :198-203
            draws[-1] = (draws[-1][0], draws[-1][1] + 1, draws[-1][2])
        else:
            draws.append((sp.atlas, 1, len(verts) // 4 - 1))
    return np.array(verts, dtype=abi.BILLBOARD_VERTEX_DTYPE), draws


def random_frame(seed, n, atlases=3, empty_crops=True):
    """Seeded synthetic frame: n billboards over a few atlases, rotated / scaled transforms, depth ties and
    signed zeros, out-of-range colours, some null sprites and some empty crops."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        ang = rng.uniform(0, 6.283)
        s = rng.uniform(0.2, 3.0)
        ca, sa = np.cos(ang) * s, np.sin(ang) * s
        tr = (ca, sa, rng.uniform(-0.2, 0.2), -sa, ca, rng.uniform(-0.2, 0.2), 0, 0, 1, rng.uniform(-20, 20), rng.uniform(-5, 5), rng.uniform(-20, 20))
        depth = float(rng.choice([rng.uniform(-50, 50), rng.integers(-3, 4), 0.0, -0.0, 7.25]))
        crop_min, crop_max = (0.0, 0.0), (1.0, 1.0)
        r = rng.random()
        if r < 0.25:
            crop_min = (rng.uniform(0, 0.6), rng.uniform(0, 0.6)); crop_max = (crop_min[0] + rng.uniform(0.05, 0.4), crop_min[1] + rng.uniform(0.05, 0.4))
        elif r < 0.32 and empty_crops:
            crop_min, crop_max = (0.7, 0.1), (0.7, 0.9)          # empty in x
        atlas = 0 if rng.random() < 0.03 else int(rng.integers(1, atlases + 1)) * 1000 + 7
        out.append(abi.make_billboard(
            atlas=atlas, min_vert=(rng.uniform(0, 0.2), rng.uniform(0, 0.2)), max_vert=(rng.uniform(0.8, 1.0), rng.uniform(0.8, 1.0)),
            uv_scale=(rng.uniform(0.05, 0.3), rng.uniform(0.05, 0.3)), uv_bias=(rng.uniform(0, 0.7), rng.uniform(0, 0.7)),
            transform=tr, color=tuple(rng.uniform(-0.3, 1.4, size=4)), depth=depth,
            anchor=(rng.uniform(0, 1), rng.uniform(0, 1)), crop_min=crop_min, crop_max=crop_max))
    return out
