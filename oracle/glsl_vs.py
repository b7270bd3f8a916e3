"""TEST INFRASTRUCTURE ONLY: ctypes loader for oracle/_ref/libspear_ref_glslvs.so -- the reference's SHIPPED particle
vertex shader (Particle_vs_source_glsl330, /root/reference/src/game/shader/Particle.h:239) executed on the CPU through
oracle/glsl/glsl_shim.h (built by oracle/glsl/build_glsl_vs.py). It is the pin of oracle/vertex_stage.py and of the
CUDA vertex stage (row A12 / N2). Never imported by the product package (spear_b200)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "libspear_ref_glslvs.so")

# ParticleSystem.cpp:253-258: 64 samples per curve, 8 x 256 types, 2 rows (spline, gradient) per type
SPLINE_SAMPLES, ATLAS_TYPES_X, ATLAS_TYPES_Y, ROWS_PER_TYPE = 64, 8, 256, 2
ATLAS_W, ATLAS_H = SPLINE_SAMPLES * ATLAS_TYPES_X, ATLAS_TYPES_Y * ROWS_PER_TYPE

_lib = None


def have():
    return os.path.isfile(PATH)


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(PATH)
        lib.glslvs_source.restype = C.c_char_p
        lib.glslvs_shade.restype = C.c_int
        lib.glslvs_shade.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                     C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib = lib
    return _lib


def source():
    return _load().glslvs_source().decode("ascii")


def atlas_from_luts(luts):
    """luts: {typeId: 512 bytes (row 0 spline, row 1 gradient, 64 RGBA8 texels each)} -> the 512 x 512 RGBA8 atlas the
    reference fills with sg_bqq_copy_subimages (ParticleSystem.cpp:323-324, :403-404, :421-449): type t sits at
    x = (t / 256) * 64, y = (t % 256) * 2."""
    atlas = np.zeros((ATLAS_H, ATLAS_W, 4), dtype=np.uint8)
    for type_id, lut in luts.items():
        rows = np.asarray(lut, dtype=np.uint8).reshape(ROWS_PER_TYPE, SPLINE_SAMPLES, 4)
        x = (type_id // ATLAS_TYPES_Y) * SPLINE_SAMPLES
        y = (type_id % ATLAS_TYPES_Y) * ROWS_PER_TYPE
        atlas[y:y + ROWS_PER_TYPE, x:x + SPLINE_SAMPLES] = rows
    return atlas


def shade(particles, instance_ubo, type_ubo, atlas, fog=None):
    """Same contract as oracle.vertex_stage.shade, except that the curves come from the ATLAS through the type UBO's
    u_SplineMad (as on the GPU of the reference) instead of from the type's own two rows. fog: uint8 [h, w, 2] or None."""
    p = np.ascontiguousarray(np.ascontiguousarray(particles).view(np.float32).reshape(-1, 8))
    n = p.shape[0]
    inst = np.ascontiguousarray(np.asarray(instance_ubo, dtype=np.float32).reshape(-1)[:40])
    typ = np.ascontiguousarray(np.asarray(type_ubo, dtype=np.float32).reshape(-1)[:24])
    atlas = np.ascontiguousarray(atlas, dtype=np.uint8)
    assert atlas.ndim == 3 and atlas.shape[2] == 4
    out = np.zeros((4 * n, 16), dtype=np.float32)
    fog_ptr, fw, fh = None, 0, 0
    if fog is not None:
        fog = np.ascontiguousarray(fog, dtype=np.uint8)
        assert fog.ndim == 3 and fog.shape[2] == 2
        fog_ptr, fw, fh = fog.ctypes.data, fog.shape[1], fog.shape[0]
    rc = _load().glslvs_shade(p.ctypes.data, n, inst.ctypes.data, typ.ctypes.data, atlas.ctypes.data, atlas.shape[1], atlas.shape[0],
                              fog_ptr, fw, fh, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("glslvs_shade failed (%d)" % rc)
    return out
