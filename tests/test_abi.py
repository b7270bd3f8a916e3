"""CPU tests of the boundary: the C-ABI library loads here (no GPU), exports every symbol
include/spear_particles.h declares, fails loudly without a device, and its host-only logic
(type bake) matches the golden vectors generated from the reference."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import scenarios as S
from spear_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "spear_particles.h")).read()
    return sorted(set(re.findall(r"SPR_API\s+[\w\s\*]+?\b(spr\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def particles():
    from spear_b200 import build
    build.build()
    from spear_b200 import particles as p
    return p


def test_library_exports_every_declared_symbol(particles):
    syms = declared_symbols()
    assert len(syms) >= 30
    lib = C.CDLL(particles.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(particles.EXPORTS) == syms


def test_struct_sizes_match_reference_layouts():
    assert C.sizeof(abi.GpuParticle) == 32                # ParticleSystem.cpp:178-184
    assert C.sizeof(abi.VertexTypeUbo) == 96              # src/game/shader/Particle.h:112-125
    assert C.sizeof(abi.VertexInstanceUbo) == 160         # src/game/shader/Particle.h:100-108
    assert C.sizeof(abi.RandomVec3) == 72                 # sv::RandomVec3 = 3 Vec3 + 36 B RandomSphere
    assert C.sizeof(abi.Mat34) == 48 and C.sizeof(abi.Frustum) == 128


def test_no_cpu_fallback(particles):
    """Without a CUDA device the product refuses to create a context (it never falls back to a CPU path)."""
    from conftest import HAVE_GPU
    if HAVE_GPU:
        pytest.skip("a GPU is present")
    with pytest.raises(particles.SpearError) as ei:
        particles.ParticleContext()
    assert "no CUDA device" in str(ei.value)


def test_product_does_not_reference_oracle():
    """Nothing under spear_b200/ or include/ may import, link or name anything under oracle/."""
    bad = []
    for base in ("spear_b200", "include"):
        for dirpath, _dirs, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if not f.endswith((".py", ".cpp", ".cu", ".h")):
                    continue
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"libspear_oracle|libspear_ref|from oracle|import oracle|oracle_api\.h|orc_", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_component_defaults_match_python_mirror(particles):
    c = abi.ParticleSystemComponent()
    particles.lib.sprComponentDefaults(C.byref(c))
    d = abi.default_component()
    assert bytes(c) == bytes(d)


def test_host_type_bake_matches_golden(particles):
    """initEffectType (ParticleSystem.cpp:298-450) is host work in the reference too: LUT + type UBO."""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "types.json")))
    comps = dict(c1=S.comp_c1(), fountain=S.comp_fountain(), c2_0=S.comp_c2(0), c2_1=S.comp_c2(1), c2_2=S.comp_c2(2),
                 c2_3=S.comp_c2(3), defaults=abi.make_component())
    for name, comp in comps.items():
        lut, ubo = particles.bake_effect_type(comp, g[name]["typeId"])
        assert lut.tobytes().hex() == g[name]["lut"], name
        assert bytes(ubo).hex() == g[name]["ubo"], name


def test_bake_many_type_ids(particles):
    """Atlas addressing (:323-324, :403-413) for type ids across atlas columns."""
    comp = S.comp_c1()
    for t in (0, 1, 255, 256, 257, 2047):
        _, ubo = particles.bake_effect_type(comp, t)
        assert ubo.u_SplineMad.x == 63.0 / 512.0 and ubo.u_SplineMad.y == 1.0 / 512.0
        assert ubo.u_SplineMad.z == ((t // 256) * 64 + 0.5) / 512.0
        assert ubo.u_SplineMad.w == ((t % 256) * 2 + 0.5) / 512.0


def _round_f32(fr):
    """the float32 nearest to the exact rational `fr` (ties to even) -- one rounding, as an FMA does"""
    from fractions import Fraction
    c = np.float32(float(fr))
    best = None
    for cand in (np.nextafter(c, np.float32(-np.inf)), c, np.nextafter(c, np.float32(np.inf))):
        err = abs(Fraction(float(cand)) - fr)
        even = (int(np.float32(cand).view(np.uint32)) & 1) == 0
        if best is None or err < best[0] or (err == best[0] and even):
            best = (err, cand)
    return np.float32(best[1])


def test_unorm8_division_free():
    """spear_b200/csrc/vertex_stage.cu unorm8(): q = v * (1/255), then fma(fma(-q, 255, v), 1/255, q) must be the
    IEEE quotient v / 255 for every 8-bit v (what GL's UNORM8 -> float conversion and oracle/glsl/glsl_shim.h compute)."""
    from fractions import Fraction
    r = np.float32(1.0) / np.float32(255.0)
    for v in range(256):
        x = np.float32(v)
        q = np.float32(x * r)
        rem = _round_f32(Fraction(float(x)) - Fraction(float(q)) * 255)
        got = _round_f32(Fraction(float(rem)) * Fraction(float(r)) + Fraction(float(q)))
        assert got == np.float32(x / np.float32(255.0)), v
