#!/usr/bin/env python3
"""Generate tests/golden/vertex_stage.json from the reference's SHIPPED particle vertex shader executed on the CPU
(oracle/_ref/libspear_ref_glslvs.so = Particle_vs_source_glsl330 of /root/reference/src/game/shader/Particle.h:239
compiled against oracle/glsl/glsl_shim.h by oracle/glsl/build_glsl_vs.py). Row A12 / N2 of SURVEY.md section 8.

Each case stores its inputs (particle records, instance and type UBO, the type's 512-byte LUT and its type id, the
RG8 fog image or none) and the 4n x 16 floats the shader wrote, all as little-endian hex, so that the numpy
restatement (oracle/vertex_stage.py) stays checked where oracle/_ref is absent. Run in the container that has
/root/reference:  python tests/golden/make_golden_vertex_stage.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "glsl"))
from oracle import glsl_vs as G  # noqa: E402

F = np.float32


def type_ubo(type_id, frame_count, frame_rate, rot, start_frame, scale, life, rel):
    """Particle_VertexType_t as ParticleSystem.cpp:312-321, :406-419 fills it for type `type_id`."""
    t = np.zeros(24, dtype=F)
    t[0:2] = frame_count; t[2] = frame_rate; t[4:8] = rot; t[9] = start_frame
    ax, ay = type_id // 256, type_id % 256
    t[12] = F(63) / F(512); t[13] = F(1) / F(512)
    t[14] = (F(ax * 64) + F(0.5)) / F(512); t[15] = (F(ay * 2) + F(0.5)) / F(512)
    t[16:18] = scale; t[18] = life[0]; t[19] = F(life[1]) * F(1.0 / 16777216.0); t[20] = rel
    return t


def smooth_lut(rng):
    """curves like the baked ones: smooth ramps with a little noise, every channel used"""
    x = np.linspace(0, 1, 64)
    lut = np.zeros((2, 64, 4))
    for r in range(2):
        for c in range(4):
            a, b, ph = rng.uniform(0.2, 1.0), rng.uniform(0.5, 3.0), rng.uniform(0, 6.28)
            lut[r, :, c] = np.clip(0.5 + 0.5 * a * np.sin(b * 3.14159 * x + ph), 0, 1)
    lut[0, :, 3] = np.clip(lut[0, :, 3] * 1.3, 0, 1)   # erosion reaches 1 (erode = 1 branch) and small values
    lut[0, 40:44, 3] = 0.0                              # below the 0.01 floor of :130
    return np.round(lut * 255).astype(np.uint8).reshape(-1)


def cases():
    rng = np.random.default_rng(20261020)
    out = []
    for k, (type_id, with_fog, rough) in enumerate([(0, False, False), (5, True, False), (300, True, False), (2047, False, False), (777, True, True)]):
        n = 48
        lut = rng.integers(0, 256, size=512, dtype=np.uint8) if rough else smooth_lut(rng)
        p = np.zeros((n, 8), dtype=F)
        p[:, 0:3] = rng.normal(0, 5, (n, 3)); p[:, 3] = rng.uniform(-0.02, 1.0, n)
        p[:, 4:7] = rng.normal(0, 2, (n, 3)); p[:, 7] = np.floor(rng.uniform(0, 1 << 24, n))
        p[0, 3] = 1.0; p[1, 3] = 0.0; p[2, 3] = -0.5; p[3, 3] = 0.5   # lifeTime 0, 1, beyond the end, the middle
        ang = rng.uniform(0, 3.0)
        inst = np.zeros(40, dtype=F)
        inst[0:16] = np.array([np.cos(ang), 0, np.sin(ang), 0, 0, 1.2, 0, 0, -np.sin(ang), 0, np.cos(ang), 0, 3, 1, -2, 1], dtype=F)
        inst[16:32] = rng.normal(0, 1, 16).astype(F)
        inst[32:36] = (0.02, 0.03, 0.4, 0.3); inst[36] = F(0.004 * k); inst[37] = F(0.5625)
        typ = type_ubo(type_id, (4, 2) if k % 2 == 0 else (3, 3), 12.0 - k, rng.uniform(-3, 3, 4), float(k % 2), (0.6, 0.4), (0.8, 0.5), float(k % 2) * 0.7)
        fog = rng.integers(0, 256, size=(6, 8, 2), dtype=np.uint8) if with_fog else None
        atlas = G.atlas_from_luts({type_id: lut, (type_id + 1) % 2048: rng.integers(0, 256, size=512, dtype=np.uint8),
                                   (type_id + 256) % 2048: rng.integers(0, 256, size=512, dtype=np.uint8)})
        got = G.shade(p, inst, typ, atlas, fog)
        out.append(dict(name="type%d_%s%s" % (type_id, "fog" if with_fog else "nofog", "_rough" if rough else ""), type_id=type_id, num_particles=n,
                        particles=p.tobytes().hex(), instance_ubo=inst.tobytes().hex(), type_ubo=typ.tobytes().hex(), lut=lut.tobytes().hex(),
                        fog=None if fog is None else dict(width=8, height=6, texels=fog.tobytes().hex()),
                        vertices=got.tobytes().hex()))
    return out


if __name__ == "__main__":
    src = G.source()
    doc = dict(
        source="Particle_vs_source_glsl330, /root/reference/src/game/shader/Particle.h:239, executed through oracle/glsl/glsl_shim.h",
        shader_sha256=hashlib.sha256(src.encode("ascii")).hexdigest(), shader_bytes=len(src),
        layout="vertices: 4 * num_particles x 16 float32 {gl_Position, v_Uv0, v_Uv1, v_Color, v_Params}", cases=cases())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vertex_stage.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")
