"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's particle vertex shader
(`/root/reference/src/game/shader/Particle.glsl:73-140`, `vs main`), SURVEY.md 8(f) row N2.

Parity status: PINNED against the reference's own shader text. There is no GL driver in this environment
(profiles/r02_probe_gl.md), so the GLSL 3.30 source the reference ships (Particle_vs_source_glsl330,
src/game/shader/Particle.h:239) is executed on the CPU instead: oracle/glsl/build_glsl_vs.py decodes it out of the
reference tree and compiles its body unchanged against a small GLSL-semantics header (oracle/glsl/glsl_shim.h) into
oracle/_ref/libspear_ref_glslvs.so. tests/test_vertex_stage_oracle.py compares this restatement with that library
(and with tests/golden/vertex_stage.json made from it) plus the hand-computed known answers: every output is
bit-identical except where cos / sin / pow enter (numpy's float32 kernels against libm: a few ulp).
What the shader leaves to the implementation is fixed the same way here, in the shim and in the CUDA kernel:
  * mat4 * vec4 is evaluated as ((c0*x + c1*y) + c2*z) + c3*w;
  * `texture()` with SG_FILTER_LINEAR (ParticleSystem.cpp:289-290) is the weighted sum of OpenGL 3.3 section 3.8.11
    with exact float weights at the uv the shader computes from u_SplineMad, UNORM8 texel = c / 255 (graphics
    hardware quantises the weights to 8 fractional bits: up to 1/256 of the difference of two adjacent texels).

Never imported by the product package (spear_b200); used by tests/ only.
"""
import numpy as np

F = np.float32


def _fract(x):
    return x - np.floor(x)


def _mod(x, y):
    return x - y * np.floor(x / y)


def _clamp01(x):
    return np.minimum(np.maximum(x, F(0)), F(1))


def srgb_to_linear(x):
    """Particle.glsl:59-66"""
    x = _clamp01(x)
    lo = F(0.0773993808) * x
    hi = np.power(x * F(0.947867298578) + F(0.0521327014218), F(2.4)).astype(F)
    return np.where(x <= F(0.04045), lo, hi).astype(F)


ATLAS_W, ATLAS_H = 512, 512  # ParticleSystem.cpp:253-258: 64 samples x 8 cells, 256 cells x 2 rows


def _unorm(texels):
    return (np.asarray(texels, dtype=np.uint8).astype(F) / F(255.0)).astype(F)


def _weights(a, b):
    """OpenGL 3.3 section 3.8.11, LINEAR: the four weights of the fractional parts (a, b) of the texel coordinate."""
    one = F(1)
    return ((one - a) * (one - b)).astype(F), (a * (one - b)).astype(F), ((one - a) * b).astype(F), (a * b).astype(F)


def _bilinear(w, t00, t10, t01, t11):
    return (((t00 * w[0] + t10 * w[1]) + t01 * w[2]) + t11 * w[3]).astype(F)


def _fog_fetch(fog, u, v):
    """textureLod(u_VisFogTexture, uv, 0).xy: RG8, LINEAR, clamp to edge. fog: uint8 array [h, w, 2]."""
    h, w = fog.shape[:2]
    tex = _unorm(fog)
    x = (u * F(w) - F(0.5)).astype(F)
    y = (v * F(h) - F(0.5)).astype(F)
    fx0, fy0 = np.floor(x), np.floor(y)
    wts = _weights((x - fx0).astype(F), (y - fy0).astype(F))
    x0 = fx0.astype(np.int64); y0 = fy0.astype(np.int64)
    x1, y1 = x0 + 1, y0 + 1
    x0 = np.clip(x0, 0, w - 1); x1 = np.clip(x1, 0, w - 1)
    y0 = np.clip(y0, 0, h - 1); y1 = np.clip(y1, 0, h - 1)
    out = []
    for k in range(2):
        out.append(_bilinear(wts, tex[y0, x0, k], tex[y0, x1, k], tex[y1, x0, k], tex[y1, x1, k]))
    return out[0], out[1]


def shade(particles, instance_ubo, type_ubo, lut_bytes, fog=None):
    """particles: structured array or [n, 8] float32 {pos.xyz, life, vel.xyz, seed} (one record per particle);
    instance_ubo: 40 floats (Particle_VertexInstance_t), type_ubo: 24 floats (Particle_VertexType_t, std140),
    lut_bytes: the type's 512-byte LUT (row 0 spline, row 1 gradient). Returns [4n, 16] float32 in
    SprShadedVertex layout {gl_Position, v_Uv0, v_Uv1, v_Color, v_Params}."""
    p = np.ascontiguousarray(particles).view(F).reshape(-1, 8)
    n = p.shape[0]
    I = np.asarray(instance_ubo, dtype=F).reshape(40)
    T = np.asarray(type_ubo, dtype=F).reshape(24)
    M, W = I[0:16], I[16:32]
    fog_mad, inv_delta, aspect = I[32:36], I[36], I[37]
    frame_count = (T[0], T[1]); frame_rate_u = T[2]
    rc = T[4:8]; start_frame = T[9]
    spline_mad = T[12:16]
    scale_base, scale_var, life_base, life_var, rel_frame_rate = T[16], T[17], T[18], T[19], T[20]
    pos, life, vel, seedp = p[:, 0:3], p[:, 3], p[:, 4:7], p[:, 7]

    def mat_point(m, v, w):  # ((c0*x + c1*y) + c2*z) + c3*w, rows 0..3
        rows = []
        for r in range(4):
            acc = (m[r] * v[:, 0] + m[4 + r] * v[:, 1]) + m[8 + r] * v[:, 2]
            if w:
                acc = acc + m[12 + r]
            rows.append(acc.astype(F))
        return rows

    # :78-80
    wp = mat_point(M, pos, True)[:3]
    wv = mat_point(M, vel, False)[:3]
    # :82-87
    abs_total = life_base + life_var * seedp
    life_left = life + inv_delta / abs_total
    life_time = F(1) - life_left
    abs_life = life_time * abs_total
    s = [_fract(np.floor(seedp * F(k)) * F(1.0 / 64.0)).astype(F) for k in (1.0, 1.0 / 64.0, 1.0 / 4096.0, 1.0 / 262144.0)]
    # :89
    wp = [(wp[i] - wv[i] * inv_delta).astype(F) for i in range(3)]
    # :94-99 texture(u_SplineTexture, splineUv) and the gradient one row below: LINEAR over the 512 x 512 atlas at
    # the uv the shader computes. Only the type's own cell (2 rows x 64 texels, `lut_bytes`) is held here: rows and
    # columns are taken relative to the cell, whose first texel centre is u_SplineMad.zw (ParticleSystem.cpp:406-413),
    # and clamped to it -- a neighbour outside the cell only ever has weight 0.
    rows = _unorm(lut_bytes).reshape(2, 64, 4)
    cell_x = int(np.floor(spline_mad[2] * F(ATLAS_W))); cell_y = int(np.floor(spline_mad[3] * F(ATLAS_H)))
    su = (_clamp01(life_time) * spline_mad[0] + spline_mad[2]).astype(F)
    sv = F(F(0) * spline_mad[1] + spline_mad[3])
    gv = F(sv + spline_mad[1])
    sx = (su * F(ATLAS_W) - F(0.5)).astype(F)
    sy = F(sv * F(ATLAS_H) - F(0.5)); gy = F(gv * F(ATLAS_H) - F(0.5))
    sfx = np.floor(sx); sfy = np.floor(sy); gfy = np.floor(gy)
    c0 = np.clip(sfx.astype(np.int64) - cell_x, 0, 63); c1 = np.clip(sfx.astype(np.int64) + 1 - cell_x, 0, 63)
    a = (sx - sfx).astype(F)

    def fetch(y, fy):
        r0 = min(max(int(fy) - cell_y, 0), 1); r1 = min(max(int(fy) + 1 - cell_y, 0), 1)
        w = _weights(a, np.full_like(a, F(y - fy)))
        return [_bilinear(w, rows[r0, c0, c], rows[r0, c1, c], rows[r1, c0, c], rows[r1, c1, c]) for c in range(4)]

    spline = fetch(sy, sfy)
    grad = fetch(gy, gfy)[:3]
    col = [srgb_to_linear(g) for g in grad]
    # :101-107
    scale = scale_base + s[0] * scale_var
    scale = scale * spline[0]
    scale = np.maximum(scale, F(0)).astype(F)
    rot_base = rc[0] + rc[1] * (s[2] * F(2) - F(1))
    rot_spin = rc[2] + rc[3] * (s[3] * F(2) - F(1))
    rotation = (rot_base + rot_spin * abs_life).astype(F)
    axx, axy = np.cos(rotation).astype(F), np.sin(rotation).astype(F)
    ayx, ayy = axy, -axx
    # :109
    ndc = mat_point(W, np.stack(wp, axis=1), True)
    # :117-125
    frame_rate = abs_life * (F(1) - rel_frame_rate) + life_time * rel_frame_rate
    frame = frame_rate * frame_rate_u + np.floor(s[0] * F(64) + s[2] * F(4096)) * start_frame
    frame_i = np.floor(frame)
    frame_d = (frame - frame_i).astype(F)
    f0 = (_mod(frame_i, frame_count[0]), _mod(np.floor(frame_i / frame_count[0]), frame_count[1]))
    f1 = (_mod(frame_i + F(1), frame_count[0]), _mod(np.floor((frame_i + F(1)) / frame_count[0]), frame_count[1]))
    # :127, :43-50
    if fog is not None:
        fx, fy = _fog_fetch(np.asarray(fog, dtype=np.uint8), (wp[0] * fog_mad[0] + fog_mad[2]).astype(F), (wp[2] * fog_mad[1] + fog_mad[3]).astype(F))
        t = _clamp01((fx - F(0.45)) * F(20))
        m = (fy * fy * t).astype(F)
        col = [(c * m).astype(F) for c in col]
    erode = np.where(spline[3] < F(1), F(1) / np.maximum(spline[3], F(0.01)), F(1)).astype(F)  # :130
    dead = life_left <= F(0)                                                                       # :113-115

    out = np.zeros((n, 4, 16), dtype=F)
    for corner in range(4):
        u, v = F(corner & 1), F((corner >> 1) & 1)                    # :91
        ox, oy = (u * F(2) - F(1)) * scale, (v * F(2) - F(1)) * scale  # :92, :103
        px = ndc[0] + (ox * axx + oy * ayx) * (F(0.5) * aspect)        # :111
        py = ndc[1] + (ox * axy + oy * ayy) * F(0.5)
        out[:, corner, 0] = np.where(dead, F(0), px)
        out[:, corner, 1] = np.where(dead, F(0), py)
        out[:, corner, 2] = np.where(dead, F(0), ndc[2])
        out[:, corner, 3] = np.where(dead, F(0), ndc[3])
        out[:, corner, 4] = (f0[0] + u) / frame_count[0]               # :133-134
        out[:, corner, 5] = (f0[1] + v) / frame_count[1]
        out[:, corner, 6] = (f1[0] + u) / frame_count[0]
        out[:, corner, 7] = (f1[1] + v) / frame_count[1]
        out[:, corner, 8], out[:, corner, 9], out[:, corner, 10] = col[0], col[1], col[2]
        out[:, corner, 11] = spline[1]                                 # :115, :129
        out[:, corner, 12] = frame_d                                   # :135-138
        out[:, corner, 13] = spline[2]
        out[:, corner, 14] = erode
        out[:, corner, 15] = F(0)
    return out.reshape(n * 4, 16)
