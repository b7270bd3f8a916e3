"""TEST INFRASTRUCTURE ONLY: ctypes loader for the CPU oracles (oracle_api.h).

kinds:
  "reference"         oracle/_ref/libspear_ref.so         (reference's own code, SSE Float4)
  "reference_scalar"  oracle/_ref/libspear_ref_scalar.so  (same, SF_FLOAT4_FORCE_SCALAR)
  "port"              oracle/libspear_oracle.so           (plain-C restatement, oracle/spear_oracle.c)

Never imported by the product package (spear_b200); used by tests/, the smoke
check and bench.py's CPU-baseline legs only.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from spear_b200 import abi  # noqa: E402

_PATHS = {
    "reference": os.path.join(HERE, "_ref", "libspear_ref.so"),
    "reference_scalar": os.path.join(HERE, "_ref", "libspear_ref_scalar.so"),
    "port": os.path.join(HERE, "libspear_oracle.so"),
    # not an oracle: integration/ParticleSystemB200.cpp (cl::ParticleSystem over the product's C ABI) behind the same
    # harness, built against the reference headers -- the thing under test in tests/test_gpu_adapter.py
    "adapter": os.path.join(HERE, "_ref", "libspear_adapter.so"),
}

GPU_PARTICLE_DTYPE = np.dtype([("pos", "<f4", 3), ("life", "<f4"), ("vel", "<f4", 3), ("seed", "<f4")])



class DrawCapture(C.Structure):
    """OrcDrawCapture (oracle_api.h): the bindings applied for one sg_draw of the last orc_render."""
    _fields_ = [("vertexBufferRing", C.c_uint32), ("vertexOffset", C.c_uint32), ("indexBuffer", C.c_uint32), ("vsSplineImage", C.c_uint32),
                ("vsVisFogImage", C.c_uint32), ("fsTextureImage", C.c_uint32), ("numElements", C.c_uint32), ("_pad", C.c_uint32)]


class AppendCapture(C.Structure):
    """OrcAppendCapture: one sg_append_buffer call of the last orc_render."""
    _fields_ = [("bufferRing", C.c_uint32), ("offset", C.c_uint32), ("numBytes", C.c_uint32), ("_pad", C.c_uint32), ("fnv1a64", C.c_uint64)]


_libs = {}


def lib_path(kind):
    return _PATHS[kind]


def have(kind):
    return os.path.isfile(_PATHS[kind])


def _load(kind):
    if kind in _libs:
        return _libs[kind]
    lib = C.CDLL(_PATHS[kind])
    vp = C.c_void_p
    u32 = C.c_uint32
    lib.orc_kind.restype = C.c_char_p
    lib.orc_create.restype = vp
    lib.orc_create.argtypes = [C.POINTER(abi.SystemsDesc)]
    lib.orc_destroy.argtypes = [vp]
    lib.orc_reserve_type.restype = u32
    lib.orc_reserve_type.argtypes = [vp, C.POINTER(abi.ParticleSystemComponent)]
    lib.orc_release_type.argtypes = [vp, C.POINTER(abi.ParticleSystemComponent)]
    lib.orc_add_effect.restype = u32
    lib.orc_add_effect.argtypes = [vp, u32, C.c_uint8, C.POINTER(abi.ParticleSystemComponent), C.POINTER(abi.Transform)]
    lib.orc_update_transform.argtypes = [vp, u32, C.POINTER(abi.TransformUpdate)]
    lib.orc_prepare_for_remove.restype = C.c_int
    lib.orc_prepare_for_remove.argtypes = [vp, u32, C.POINTER(abi.FrameArgs)]
    lib.orc_remove.argtypes = [vp, u32]
    lib.orc_update.argtypes = [vp, C.POINTER(u32), u32, C.POINTER(abi.FrameArgs)]
    lib.orc_render.restype = u32
    lib.orc_render.argtypes = [vp, C.POINTER(u32), u32, C.POINTER(abi.RenderArgs), C.POINTER(abi.DrawRecord), u32, C.POINTER(u32)]
    lib.orc_effect_info.argtypes = [vp, u32, C.POINTER(abi.EffectInfo)]
    lib.orc_read_free.restype = u32
    lib.orc_read_free.argtypes = [vp, u32, vp, u32]
    lib.orc_read_stream.restype = u32
    lib.orc_read_stream.argtypes = [vp, u32, vp, u32]
    lib.orc_read_slots.restype = u32
    lib.orc_read_slots.argtypes = [vp, u32, vp, u32]
    lib.orc_rng_state.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.orc_type_lut.argtypes = [vp, u32, vp]
    lib.orc_type_ubo.argtypes = [vp, u32, C.POINTER(abi.VertexTypeUbo)]
    lib.orc_make_frustum.argtypes = [C.POINTER(abi.Mat44), C.c_float, C.POINTER(abi.Frustum)]
    if hasattr(lib, "orc_render_capture"):   # reference and adapter harnesses only (the plain-C port has no renderer side)
        lib.orc_render_capture.restype = u32
        lib.orc_render_capture.argtypes = [vp, C.POINTER(DrawCapture), u32, C.POINTER(AppendCapture), u32, C.POINTER(u32)]
        lib.orc_read_atlas.restype = C.c_int
        lib.orc_read_atlas.argtypes = [vp, vp]
        lib.orc_area_spheres.restype = u32
        lib.orc_area_spheres.argtypes = [vp, C.POINTER(C.c_float), u32]
    lib.orc_bench_steps.restype = C.c_double
    lib.orc_bench_steps.argtypes = [vp, C.POINTER(u32), u32, C.POINTER(abi.FrameArgs), u32, C.POINTER(C.c_uint64)]
    _libs[kind] = lib
    return lib


def _ids(ids):
    arr = np.ascontiguousarray(ids, dtype=np.uint32)
    return arr, arr.ctypes.data_as(C.POINTER(C.c_uint32))


class OracleSystem:
    """One cl::ParticleSystem instance run on the CPU by an oracle library."""

    def __init__(self, kind="reference", seed=(1, 2, 3, 4)):
        self.kind = kind
        self.lib = _load(kind)
        desc = abi.make_systems_desc(seed)
        self.h = self.lib.orc_create(C.byref(desc))
        self._keep = []

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reserve_type(self, comp):
        self._keep.append(comp)
        return self.lib.orc_reserve_type(self.h, C.byref(comp))

    def release_type(self, comp):
        self.lib.orc_release_type(self.h, C.byref(comp))

    def add_effect(self, comp, transform=None, entity_id=0, component_index=0):
        self._keep.append(comp)
        t = transform if transform is not None else abi.make_transform()
        return self.lib.orc_add_effect(self.h, entity_id, component_index, C.byref(comp), C.byref(t))

    def update_transform(self, effect_id, update):
        self.lib.orc_update_transform(self.h, effect_id, C.byref(update))

    def prepare_for_remove(self, effect_id, frame_args):
        return bool(self.lib.orc_prepare_for_remove(self.h, effect_id, C.byref(frame_args)))

    def remove(self, effect_id):
        self.lib.orc_remove(self.h, effect_id)

    def update(self, ids, frame_args):
        arr, p = _ids(ids)
        self.lib.orc_update(self.h, p, len(arr), C.byref(frame_args))

    def render(self, ids, render_args, capacity=None):
        arr, p = _ids(ids)
        cap = capacity if capacity is not None else max(len(arr), 1)
        recs = (abi.DrawRecord * cap)()
        nsorted = C.c_uint32(0)
        n = self.lib.orc_render(self.h, p, len(arr), C.byref(render_args), recs, cap, C.byref(nsorted))
        return [recs[i] for i in range(n)], nsorted.value

    def render_capture(self):
        """(draws, appends) of the last render(): what went through sg_apply_bindings + sg_draw and sg_append_buffer."""
        draws = (DrawCapture * 256)()
        appends = (AppendCapture * 256)()
        na = C.c_uint32(0)
        nd = self.lib.orc_render_capture(self.h, draws, 256, appends, 256, C.byref(na))
        assert nd <= 256 and na.value <= 256
        return ([dict(ring=d.vertexBufferRing, offset=d.vertexOffset, index=d.indexBuffer, vs=[d.vsSplineImage, d.vsVisFogImage],
                      fs=d.fsTextureImage, elements=d.numElements) for d in draws[:nd]],
                [dict(ring=a.bufferRing, offset=a.offset, bytes=a.numBytes, hash="%016x" % a.fnv1a64) for a in appends[:na.value]])

    def read_atlas(self):
        """The 512 x 512 RGBA8 spline atlas as the sg_* calls left it (None if never written)."""
        out = np.zeros((512, 512, 4), dtype=np.uint8)
        return out if self.lib.orc_read_atlas(self.h, out.ctypes.data) else None

    def area_spheres(self):
        buf = (C.c_float * (5 * 256))()
        n = self.lib.orc_area_spheres(self.h, buf, 256)
        return [[float(buf[5 * i + k]) for k in range(5)] for i in range(min(n, 256))]

    def effect_info(self, effect_id):
        info = abi.EffectInfo()
        self.lib.orc_effect_info(self.h, effect_id, C.byref(info))
        return info

    def read_free(self, effect_id):
        n = self.lib.orc_read_free(self.h, effect_id, None, 0)
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self.lib.orc_read_free(self.h, effect_id, out.ctypes.data, n)
        return out[:n]

    def read_stream(self, effect_id):
        n = self.lib.orc_read_stream(self.h, effect_id, None, 0)
        out = np.zeros(max(n, 1), dtype=GPU_PARTICLE_DTYPE)
        self.lib.orc_read_stream(self.h, effect_id, out.ctypes.data, n)
        return out[:n]

    def read_slots(self, effect_id):
        n = self.lib.orc_read_slots(self.h, effect_id, None, 0)
        out = np.zeros(max(n, 1), dtype=GPU_PARTICLE_DTYPE)
        self.lib.orc_read_slots(self.h, effect_id, out.ctypes.data, n)
        return out[:n]

    def rng_state(self):
        out = (C.c_uint64 * 2)()
        self.lib.orc_rng_state(self.h, out)
        return int(out[0]), int(out[1])

    def type_lut(self, type_id):
        out = np.zeros(abi.SPLINE_LUT_BYTES, dtype=np.uint8)
        self.lib.orc_type_lut(self.h, type_id, out.ctypes.data)
        return out

    def type_ubo(self, type_id):
        out = abi.VertexTypeUbo()
        self.lib.orc_type_ubo(self.h, type_id, C.byref(out))
        return out

    def bench_steps(self, ids, frame_args, steps):
        arr, p = _ids(ids)
        ps = C.c_uint64(0)
        secs = self.lib.orc_bench_steps(self.h, p, len(arr), C.byref(frame_args), steps, C.byref(ps))
        return secs, ps.value


def make_frustum(world_to_clip, clip_near_w=0.0, kind="reference"):
    lib = _load(kind)
    m = abi.Mat44()
    for i in range(16):
        m.v[i] = float(world_to_clip[i])
    f = abi.Frustum()
    lib.orc_make_frustum(C.byref(m), C.c_float(clip_near_w), C.byref(f))
    return f


def fnv1a64(data):
    h = 0xcbf29ce484222325
    for b in bytes(data):
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h
