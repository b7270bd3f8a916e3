"""TEST INFRASTRUCTURE ONLY: the CPU oracles of SURVEY.md 8(f) row N5 (area query of the particle spheres).

  AreaOracle     ctypes loader of oracle/_ref/libspear_ref_area.so = the reference's own
                 src/client/AreaSystem.cpp behind oracle/ref_area_harness.cpp (spatial tree, optimize(), queryFrustum)
  port_query()   numpy float32 restatement of sf::Frustum::intersects(const sf::Sphere&) (src/sf/Frustum.h:42-58)
                 over a flat list of spheres; the reference's tree only prunes with box tests (AreaSystem.cpp:166,
                 :645) that are conservative for spheres inside the slab AreaMinY..AreaMaxY it clamps its boxes to
                 (:12-13, :347-360), so there the SET it returns is the set of spheres that pass this test; a sphere
                 reaching outside the slab can additionally be dropped by a clamped box (history dependent, no
                 functional form), so in general reference set <= this set.
Parity status: PINNED -- tests/test_area_oracle.py checks the port against the reference build on seeded random
scenes (adds, moves, removes, optimize passes), and both against tests/golden/areas.json (generated from the
reference build by tests/golden/make_golden_areas.py; the reference tree does not exist on the GPU box).

Never imported by the product package (spear_b200); used by tests/ only.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from spear_b200 import abi  # noqa: E402

LIB = os.path.join(HERE, "_ref", "libspear_ref_area.so")
F = np.float32

GROUP_PARTICLE_EFFECT = None  # filled from the reference enum by the harness (orc_area_particle_group)


def have_reference():
    return os.path.isfile(LIB)


_lib = None


def _load():
    global _lib, GROUP_PARTICLE_EFFECT
    if _lib is None:
        lib = C.CDLL(LIB)
        vp, u32 = C.c_void_p, C.c_uint32
        fp = C.POINTER(C.c_float)
        lib.orc_area_create.restype = vp
        lib.orc_area_destroy.argtypes = [vp]
        lib.orc_area_add_sphere.restype = u32
        lib.orc_area_add_sphere.argtypes = [vp, u32, u32, fp, C.c_float]
        lib.orc_area_update_sphere.argtypes = [vp, u32, fp, C.c_float]
        lib.orc_area_remove_sphere.argtypes = [vp, u32]
        lib.orc_area_optimize.argtypes = [vp]
        lib.orc_area_query.restype = u32
        lib.orc_area_query.argtypes = [vp, C.POINTER(abi.Frustum), C.POINTER(u32), u32]
        lib.orc_area_particle_group.restype = u32
        lib.orc_area_other_group.restype = u32
        GROUP_PARTICLE_EFFECT = lib.orc_area_particle_group()
        _lib = lib
    return _lib


def _vec3(v):
    return (C.c_float * 3)(float(v[0]), float(v[1]), float(v[2]))


class AreaOracle:
    """The reference's cl::AreaSystem holding sphere areas, as the particle system uses it
    (ParticleSystem.cpp:762-763 add, :782-783 update, :801 remove)."""

    def __init__(self):
        self.lib = _load()
        self.h = self.lib.orc_area_create()

    def close(self):
        if self.h:
            self.lib.orc_area_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add(self, user_id, origin, radius, group=None):
        g = GROUP_PARTICLE_EFFECT if group is None else group
        return self.lib.orc_area_add_sphere(self.h, g, user_id, _vec3(origin), C.c_float(radius))

    def update(self, area_id, origin, radius):
        self.lib.orc_area_update_sphere(self.h, area_id, _vec3(origin), C.c_float(radius))

    def remove(self, area_id):
        self.lib.orc_area_remove_sphere(self.h, area_id)

    def optimize(self):
        self.lib.orc_area_optimize(self.h)

    def query(self, frustum, capacity=1 << 20):
        """userIds of the AreaGroup::ParticleEffect areas queryFrustum returns, in the reference's (tree) order."""
        out = np.zeros(capacity, dtype=np.uint32)
        n = self.lib.orc_area_query(self.h, C.byref(frustum), out.ctypes.data_as(C.POINTER(C.c_uint32)), capacity)
        assert n <= capacity
        return out[:n].copy()


def frustum_arrays(frustum):
    """(side, caps) as float32 arrays [component k][plane j] (sf::Frustum::side[k] / caps[k] are Float4 over planes)."""
    side = np.array([[frustum.side[k][j] for j in range(4)] for k in range(4)], dtype=F)
    caps = np.array([[frustum.caps[k][j] for j in range(4)] for k in range(4)], dtype=F)
    return side, caps


def port_intersects(frustum, origins, radii):
    """sf::Frustum::intersects(const sf::Sphere&), src/sf/Frustum.h:42-58, for n spheres at once; every operation is
    a separately rounded float32 multiply or add, in the reference's order (the build uses -ffp-contract=off)."""
    side, caps = frustum_arrays(frustum)
    o = np.asarray(origins, dtype=F).reshape(-1, 3)
    r = np.asarray(radii, dtype=F).reshape(-1, 1)
    so = side[3][None, :] + r                        # Float4 so = side[3] + r, co = caps[3] + r   (:45)
    co = caps[3][None, :] + r
    for k in range(3):                               # x, y, z in turn                              (:48-55)
        ok = o[:, k:k + 1]
        so = so + (side[k][None, :] * ok).astype(F)  # so += s*o
        co = co + (caps[k][None, :] * ok).astype(F)  # co += c*o
    with np.errstate(invalid="ignore"):
        m = np.where(so < co, so, co)                # so.min(co): _mm_min_ps = a < b ? a : b       (:57, Float4.h:203)
        return np.all(m > F(0), axis=1)              # allGreaterThanZero(): a NaN lane fails


def port_query(frustum, user_ids, origins, radii):
    """The user ids of the spheres that pass the test, ascending: the defined order of sprQueryEffects."""
    ids = np.asarray(user_ids, dtype=np.uint32)
    if ids.size == 0:
        return ids
    hit = port_intersects(frustum, origins, radii)
    return np.sort(ids[hit])


# ---------------------------------------------------------------------------------------------
# Seeded scenes shared by the golden generator and the tests: a list of ops over sphere areas
#   ("add", userId, origin, radius) / ("move", userId, origin, radius) / ("remove", userId) /
#   ("optimize",) / ("query", frustumIndex)
# ---------------------------------------------------------------------------------------------

AREA_MIN_Y, AREA_MAX_Y = -2.0, 8.0   # src/client/AreaSystem.cpp:12-13


def random_scene(seed, n, rounds=4, extent=60.0, in_slab=True):
    """in_slab: every sphere lies inside AreaMinY <= y <= AreaMaxY, the slab the reference clamps its tree boxes to
    (clampBounds, AreaSystem.cpp:347-360). There the tree's box tests are conservative and queryFrustum returns exactly
    the spheres that pass Frustum::intersects; a sphere reaching outside the slab can additionally be pruned by a
    clamped box, so for those scenes the reference's set is only a subset of the sphere test's."""
    rng = np.random.default_rng(seed)
    ops, live, next_id = [], [], 0

    def sphere():
        if in_slab:
            r = float(F(rng.choice([0.0, 0.25, 1.0, 2.5, 4.5])))
            y = rng.uniform(AREA_MIN_Y + r + 0.01, AREA_MAX_Y - r - 0.01)
        else:
            r = float(F(rng.choice([0.0, 0.25, 1.0, 2.5, 6.0, 20.0])))
            y = rng.uniform(-4.0, 12.0)
        o = (rng.uniform(-extent, extent), y, rng.uniform(-extent, extent))
        return tuple(float(F(x)) for x in o), r

    radius_of = {}
    for _ in range(n):
        o, r = sphere()
        ops.append(("add", next_id, o, r))
        radius_of[next_id] = r
        live.append(next_id)
        next_id += 1
    for rnd in range(rounds):
        ops.append(("query", rnd % 4))
        for _ in range(max(1, len(live) // 4)):
            if not live:
                break
            u = live[int(rng.integers(len(live)))]
            while True:   # a move keeps the radius: updateSphereArea(areaId, { position, type.updateRadius }), :782-783
                o, r = sphere()
                if r == radius_of[u]:
                    break
            ops.append(("move", u, o, r))
        for _ in range(len(live) // 8):
            k = int(rng.integers(len(live)))
            ops.append(("remove", live.pop(k)))
        for _ in range(n // 8):
            o, r = sphere()
            ops.append(("add", next_id, o, r))
            radius_of[next_id] = r
            live.append(next_id)
            next_id += 1
        if rnd % 2 == 0:
            ops.append(("optimize",))
        ops.append(("query", (rnd + 1) % 4))
    return ops


def run_scene_reference(ops, frusta):
    """The scene on the reference's AreaSystem; each query's ParticleEffect user ids, sorted ascending.
    A few areas of another group are registered next to the particle ones; they must never be returned."""
    o = AreaOracle()
    area_of = {}
    other = o.lib.orc_area_other_group()
    out = []
    try:
        for i, op in enumerate(ops):
            if op[0] == "add":
                area_of[op[1]] = o.add(op[1], op[2], op[3])
                if i % 7 == 0:
                    o.add(0x80000000 | op[1], op[2], op[3], group=other)
            elif op[0] == "move":
                o.update(area_of[op[1]], op[2], op[3])
            elif op[0] == "remove":
                o.remove(area_of.pop(op[1]))
            elif op[0] == "optimize":
                o.optimize()
            else:
                got = o.query(frusta[op[1]])
                assert len(set(got.tolist())) == len(got)
                out.append(np.sort(got))
    finally:
        o.close()
    return out


def run_scene_port(ops, frusta):
    spheres = {}
    out = []
    for op in ops:
        if op[0] in ("add", "move"):
            spheres[op[1]] = (op[2], op[3])
        elif op[0] == "remove":
            del spheres[op[1]]
        elif op[0] == "query":
            ids = list(spheres.keys())
            out.append(port_query(frusta[op[1]], ids, [spheres[u][0] for u in ids], [spheres[u][1] for u in ids]))
    return out
