"""ctypes binding of include/spear_particles.h (the reference-facing plugin API).

ParticleSystem mirrors cl::ParticleSystem (src/client/ParticleSystem.h:9-22) method for
method; ParticleContext is the per-GPU owner (device, stream, slot pool). Everything
here calls straight into libspear_particles.so -- if the library is missing this
module raises ImportError rather than falling back to anything.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspear_particles.so")
# A/B measurement aid: SPR_LIB_VARIANT=name loads spear_b200/libspear_particles.name.so (an experiment build of the same sources)
if os.environ.get("SPR_LIB_VARIANT"):
    LIB_PATH = LIB_PATH[:-3] + "." + os.environ["SPR_LIB_VARIANT"] + ".so"

GPU_PARTICLE_DTYPE = np.dtype([("pos", "<f4", 3), ("life", "<f4"), ("vel", "<f4", 3), ("seed", "<f4")])

EXPORTS = [
    "sprComponentDefaults", "sprGetLastError", "sprContextCreate", "sprContextDestroy", "sprContextSynchronize",
    "sprContextStream", "sprContextLaunchCount", "sprSystemCreate", "sprSystemDestroy", "sprReserveEffectType",
    "sprReleaseEffectType", "sprAddEffect", "sprUpdateTransform", "sprPrepareForRemove", "sprRemoveEffect",
    "sprUpdateParticles", "sprUpdateParticlesBatch", "sprRenderMain", "sprGetEffectInfo", "sprReadFreeIndices",
    "sprReadVertexStream", "sprReadSlots", "sprReadStreamSlots", "sprGetRngState", "sprGetEffectTypeLut",
    "sprGetEffectTypeUbo", "sprGetEffectDeviceView", "sprPackRuns", "sprExpandRuns",
    "sprContextEnableKernelTiming", "sprContextGetKernelTimes", "sprBakeEffectType", "sprRenderMainBatch", "sprContextGetTotals",
    "sprRunBufferCreate", "sprRunBufferDestroy", "sprRunBufferGetIpcHandle", "sprRunBufferOpenPeer", "sprRunBufferClosePeer",
    "sprRunBufferDevicePtr", "sprRunBufferPack", "sprExpandFromPeers", "sprContextSetAssemblyShare", "sprContextSetAssemblyKernel", "sprContextSetDeferredExpansion", "sprShadeVertices",
    "sprBillboardSystemCreate", "sprBillboardSystemDestroy", "sprBillboardAdd", "sprBillboardRenderMain",
    "sprBillboardReadVertices", "sprBillboardPack", "sprBillboardAddPacked",
    "sprDescriptorsFromJson", "sprDescriptorCount", "sprDescriptorGet", "sprDescriptorTextureName", "sprDescriptorsFree",
    "sprQueryEffects",
]


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "spear_b200: %s is not built (run `python spear_b200/build.py`); there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    P = C.POINTER
    lib.sprGetLastError.restype = C.c_char_p
    lib.sprComponentDefaults.argtypes = [P(abi.ParticleSystemComponent)]
    lib.sprContextCreate.argtypes = [P(abi.ContextDesc), P(vp)]
    lib.sprContextDestroy.argtypes = [vp]
    lib.sprContextSynchronize.argtypes = [vp]
    lib.sprContextStream.restype = vp
    lib.sprContextStream.argtypes = [vp]
    lib.sprContextLaunchCount.restype = u64
    lib.sprContextLaunchCount.argtypes = [vp]
    lib.sprSystemCreate.argtypes = [vp, P(abi.SystemsDesc), P(vp)]
    lib.sprSystemDestroy.argtypes = [vp]
    lib.sprReserveEffectType.argtypes = [vp, P(abi.ParticleSystemComponent), P(u32)]
    lib.sprReleaseEffectType.argtypes = [vp, P(abi.ParticleSystemComponent)]
    lib.sprAddEffect.argtypes = [vp, u32, C.c_uint8, P(abi.ParticleSystemComponent), P(abi.Transform), P(u32)]
    lib.sprUpdateTransform.argtypes = [vp, u32, P(abi.TransformUpdate)]
    lib.sprPrepareForRemove.argtypes = [vp, u32, P(abi.FrameArgs), P(C.c_int)]
    lib.sprRemoveEffect.argtypes = [vp, u32]
    lib.sprUpdateParticles.argtypes = [vp, P(u32), u32, P(abi.FrameArgs)]
    lib.sprUpdateParticlesBatch.argtypes = [vp, P(vp), u32, P(P(u32)), P(u32), P(abi.FrameArgs), C.c_size_t]
    lib.sprRenderMain.argtypes = [vp, P(u32), u32, P(abi.RenderArgs), P(abi.RenderOutput)]
    lib.sprRenderMainBatch.argtypes = [vp, P(vp), u32, P(P(u32)), P(u32), P(abi.RenderArgs), C.c_size_t, P(abi.RenderOutput)]
    lib.sprGetEffectInfo.argtypes = [vp, u32, P(abi.EffectInfo)]
    lib.sprReadFreeIndices.argtypes = [vp, u32, vp, u32, P(u32)]
    lib.sprReadVertexStream.argtypes = [vp, u32, vp, u32, P(u32)]
    lib.sprReadSlots.argtypes = [vp, u32, vp, u32, P(u32)]
    lib.sprReadStreamSlots.argtypes = [vp, u32, vp, u32, P(u32)]
    lib.sprGetRngState.argtypes = [vp, P(u64)]
    lib.sprGetEffectTypeLut.argtypes = [vp, u32, vp]
    lib.sprGetEffectTypeUbo.argtypes = [vp, u32, P(abi.VertexTypeUbo)]
    lib.sprGetEffectDeviceView.argtypes = [vp, u32, P(abi.EffectDeviceView)]
    lib.sprPackRuns.argtypes = [vp, P(vp), P(u32), u32, vp, u64, vp, P(u64)]
    lib.sprExpandRuns.argtypes = [vp, vp, u64, vp]
    lib.sprBakeEffectType.argtypes = [P(abi.ParticleSystemComponent), u32, vp, P(abi.VertexTypeUbo)]
    lib.sprContextGetTotals.argtypes = [vp, P(u64), P(u64)]
    lib.sprRunBufferCreate.argtypes = [vp, u64, P(vp)]
    lib.sprRunBufferDestroy.argtypes = [vp]
    lib.sprRunBufferGetIpcHandle.argtypes = [vp, vp]
    lib.sprRunBufferOpenPeer.argtypes = [vp, vp, P(vp)]
    lib.sprRunBufferClosePeer.argtypes = [vp, vp]
    lib.sprRunBufferDevicePtr.restype = vp
    lib.sprRunBufferDevicePtr.argtypes = [vp]
    lib.sprRunBufferPack.argtypes = [vp, P(vp), P(u32), u32, vp]
    lib.sprExpandFromPeers.argtypes = [vp, P(vp), u32, u32, vp, u64, vp]
    lib.sprContextSetAssemblyShare.argtypes = [vp, u32]
    lib.sprContextSetAssemblyKernel.argtypes = [vp, u32]
    lib.sprContextSetDeferredExpansion.argtypes = [vp, C.c_int]
    lib.sprShadeVertices.argtypes = [vp, P(abi.DrawRecord), u32, P(abi.FogImage), vp, u64, P(u64)]
    lib.sprBillboardSystemCreate.argtypes = [vp, u32, P(vp)]
    lib.sprBillboardSystemDestroy.argtypes = [vp]
    lib.sprBillboardAdd.argtypes = [vp, P(abi.Billboard), u32]
    lib.sprBillboardPack.argtypes = [P(abi.Billboard), u32, P(abi.BillboardPacked), P(u32)]
    lib.sprBillboardAddPacked.argtypes = [vp, vp, u32, C.c_int]
    lib.sprBillboardRenderMain.argtypes = [vp, P(abi.RenderArgs), P(abi.BillboardOutput)]
    lib.sprBillboardReadVertices.argtypes = [vp, vp, u32, P(u32)]
    lib.sprDescriptorsFromJson.argtypes = [C.c_char_p, C.c_size_t, P(vp)]
    lib.sprDescriptorCount.restype = u32
    lib.sprDescriptorCount.argtypes = [vp]
    lib.sprDescriptorGet.restype = P(abi.ParticleSystemComponent)
    lib.sprDescriptorGet.argtypes = [vp, u32]
    lib.sprDescriptorTextureName.restype = C.c_char_p
    lib.sprDescriptorTextureName.argtypes = [vp, u32]
    lib.sprDescriptorsFree.argtypes = [vp]
    lib.sprQueryEffects.argtypes = [vp, P(abi.Frustum), P(u32), u32, P(u32)]
    lib.sprContextEnableKernelTiming.argtypes = [vp, C.c_int]
    lib.sprContextGetKernelTimes.argtypes = [vp, C.c_int, P(C.c_char_p), P(C.c_double), P(u64), u32, P(u32)]
    return lib


lib = _load()


class SpearError(RuntimeError):
    pass


def _check(status):
    if status != 0:
        raise SpearError("spear_particles status %d: %s" % (status, lib.sprGetLastError().decode()))


def _ids(ids):
    arr = np.ascontiguousarray(ids, dtype=np.uint32)
    return arr, arr.ctypes.data_as(C.POINTER(C.c_uint32))


def bake_effect_type(comp, type_id=0):
    """Host-only LUT + type UBO bake (no GPU needed): (lut uint8[512], VertexTypeUbo)."""
    lut = np.zeros(abi.SPLINE_LUT_BYTES, dtype=np.uint8)
    ubo = abi.VertexTypeUbo()
    _check(lib.sprBakeEffectType(C.byref(comp), type_id, lut.ctypes.data, C.byref(ubo)))
    return lut, ubo


class ParticleContext:
    """One GPU: device, stream, slot pool, tables (SprContext)."""

    def __init__(self, device=0, sort_particles=False, max_particles_per_frame=0, max_cache_frames=0,
                 initial_slot_capacity=0, stream=None):
        desc = abi.ContextDesc()
        desc.device = device
        desc.flags = abi.CTX_SORT_PARTICLES if sort_particles else 0
        desc.maxParticlesPerFrame = max_particles_per_frame
        desc.maxParticleCacheFrames = max_cache_frames
        desc.initialSlotCapacity = initial_slot_capacity
        desc.stream = stream
        self.h = C.c_void_p()
        _check(lib.sprContextCreate(C.byref(desc), C.byref(self.h)))
        self.systems = []

    def close(self):
        if self.h:
            for s in list(self.systems):
                s.close()
            lib.sprContextDestroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _check(lib.sprContextSynchronize(self.h))

    @property
    def stream(self):
        return lib.sprContextStream(self.h)

    @property
    def launch_count(self):
        return int(lib.sprContextLaunchCount(self.h))

    def totals(self, alive=True):
        """(sim steps issued so far, live particles now) over every live effect of the context."""
        steps, n = C.c_uint64(), C.c_uint64()
        _check(lib.sprContextGetTotals(self.h, C.byref(steps), C.byref(n) if alive else None))
        return int(steps.value), int(n.value)

    def enable_kernel_timing(self, enable=True):
        _check(lib.sprContextEnableKernelTiming(self.h, 1 if enable else 0))

    def kernel_times(self, reset=True):
        """{kernel name: (total ms, launches)} measured with CUDA events on the context stream."""
        names = (C.c_char_p * 24)()
        ms = (C.c_double * 24)()
        launches = (C.c_uint64 * 24)()
        n = C.c_uint32()
        _check(lib.sprContextGetKernelTimes(self.h, 1 if reset else 0, names, ms, launches, 24, C.byref(n)))
        return {names[i].decode(): (ms[i], int(launches[i])) for i in range(n.value)}

    def create_system(self, seed=(1, 2, 3, 4)):
        return ParticleSystem(self, seed)

    def update_batch(self, systems, id_lists, frame_args):
        """sprUpdateParticlesBatch: one set of launches for many systems (shared FrameArgs)."""
        n = len(systems)
        handles = (C.c_void_p * n)(*[s.h for s in systems])
        arrs = [np.ascontiguousarray(ids, dtype=np.uint32) for ids in id_lists]
        ptrs = (C.POINTER(C.c_uint32) * n)(*[a.ctypes.data_as(C.POINTER(C.c_uint32)) for a in arrs])
        counts = (C.c_uint32 * n)(*[len(a) for a in arrs])
        if isinstance(frame_args, (list, tuple)):   # one FrameArgs per system
            assert len(frame_args) == n
            arr = (type(frame_args[0]) * n)(*frame_args)
            _check(lib.sprUpdateParticlesBatch(self.h, handles, n, ptrs, counts, arr, C.sizeof(type(frame_args[0]))))
            return
        _check(lib.sprUpdateParticlesBatch(self.h, handles, n, ptrs, counts, C.byref(frame_args), 0))

    def make_batch(self, systems, id_lists):
        """Pre-marshalled batch for benchmark loops (no per-step Python marshalling)."""
        n = len(systems)
        handles = (C.c_void_p * n)(*[s.h for s in systems])
        arrs = [np.ascontiguousarray(ids, dtype=np.uint32) for ids in id_lists]
        ptrs = (C.POINTER(C.c_uint32) * n)(*[a.ctypes.data_as(C.POINTER(C.c_uint32)) for a in arrs])
        counts = (C.c_uint32 * n)(*[len(a) for a in arrs])
        return (handles, ptrs, counts, n, arrs)

    def update_prepared(self, batch, frame_args):
        handles, ptrs, counts, n, _ = batch
        _check(lib.sprUpdateParticlesBatch(self.h, handles, n, ptrs, counts, C.byref(frame_args), 0))

    def render_prepared(self, batch, render_args):
        """sprRenderMainBatch over a pre-marshalled batch; returns the list of RenderOutput structs."""
        handles, ptrs, counts, n, _ = batch
        outs = (abi.RenderOutput * n)()
        _check(lib.sprRenderMainBatch(self.h, handles, n, ptrs, counts, C.byref(render_args), 0, outs))
        return outs

    def pack_runs(self, systems, effect_ids, device_out, capacity_records, device_counts=None, want_total=True):
        n = len(systems)
        handles = (C.c_void_p * n)(*[s.h for s in systems])
        arr, p = _ids(effect_ids)
        total = C.c_uint64(0)
        _check(lib.sprPackRuns(self.h, handles, p, n, device_out, capacity_records, device_counts,
                               C.byref(total) if want_total else None))
        return total.value

    # ---- fused gather + expansion over peer memory (see include/spear_particles.h)
    def run_buffer_create(self, capacity_records):
        h = C.c_void_p()
        _check(lib.sprRunBufferCreate(self.h, capacity_records, C.byref(h)))
        return h

    def run_buffer_ipc_handle(self, buf):
        out = (C.c_uint8 * 64)()
        _check(lib.sprRunBufferGetIpcHandle(buf, out))
        return bytes(out)

    def run_buffer_open_peer(self, handle_bytes):
        arr = (C.c_uint8 * 64).from_buffer_copy(handle_bytes)
        p = C.c_void_p()
        _check(lib.sprRunBufferOpenPeer(self.h, arr, C.byref(p)))
        return p.value

    def run_buffer_ptr(self, buf):
        return lib.sprRunBufferDevicePtr(buf)

    def make_pack_list(self, systems, effect_ids):
        n = len(systems)
        handles = (C.c_void_p * n)(*[s.h for s in systems])
        arr = np.ascontiguousarray(effect_ids, dtype=np.uint32)
        return (handles, arr, arr.ctypes.data_as(C.POINTER(C.c_uint32)), n)

    def run_buffer_pack(self, pack_list, buf):
        handles, _arr, p, n = pack_list
        _check(lib.sprRunBufferPack(self.h, handles, p, n, buf))

    def expand_from_peers(self, buffer_ptrs, device_vertices_out, capacity_records, device_total=None, my_rank=0):
        n = len(buffer_ptrs)
        arr = (C.c_void_p * n)(*buffer_ptrs)
        _check(lib.sprExpandFromPeers(self.h, arr, n, my_rank, device_vertices_out, capacity_records, device_total))

    def set_assembly_share(self, ctas_per_sm):
        """Resident CTAs per SM of expand_from_peers (0 = the whole GPU)."""
        _check(lib.sprContextSetAssemblyShare(self.h, ctas_per_sm))

    def set_assembly_kernel(self, kernel):
        """0 = by share, 1 = load/store kernel, 2 = bulk-asynchronous pipeline (expand_from_peers)."""
        _check(lib.sprContextSetAssemblyKernel(self.h, kernel))

    def set_deferred_expansion(self, on):
        """Sort mode: updates stop writing the local vertex stream, the pack calls gather the sorted records from the slot state."""
        _check(lib.sprContextSetDeferredExpansion(self.h, 1 if on else 0))

    def expand_runs(self, device_records, num_records, device_vertices_out):
        _check(lib.sprExpandRuns(self.h, device_records, num_records, device_vertices_out))


class DescriptorSet:
    """Effect descriptors read from the reference's JSON prefab format (sprDescriptorsFromJson; host only).
    The descriptors live as long as this object: keep it while effects created from them exist."""

    def __init__(self, json_text):
        data = json_text.encode() if isinstance(json_text, str) else bytes(json_text)
        self.h = C.c_void_p()
        _check(lib.sprDescriptorsFromJson(data, len(data), C.byref(self.h)))

    def __len__(self):
        return lib.sprDescriptorCount(self.h)

    def __getitem__(self, i):
        if not 0 <= i < len(self):
            raise IndexError(i)
        return lib.sprDescriptorGet(self.h, i).contents

    def texture_name(self, i):
        return lib.sprDescriptorTextureName(self.h, i).decode()

    def close(self):
        if self.h:
            lib.sprDescriptorsFree(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BillboardSystem:
    """cl::BillboardSystem over the C ABI (same calls as oracle.orc.BillboardOracle)."""

    def __init__(self, ctx, max_billboards_per_frame=0):
        self.ctx = ctx
        self.h = C.c_void_p()
        _check(lib.sprBillboardSystemCreate(ctx.h, max_billboards_per_frame, C.byref(self.h)))

    def close(self):
        if self.h:
            lib.sprBillboardSystemDestroy(self.h)
            self.h = None

    def add(self, billboards):
        n = len(billboards)
        arr = (abi.Billboard * max(n, 1))(*billboards)
        _check(lib.sprBillboardAdd(self.h, arr, n))

    @staticmethod
    def pack(billboards):
        """sprBillboardPack: the ctypes array of SprBillboardPacked addBillboard would make of these (unloaded sprites dropped)."""
        n = len(billboards)
        arr = (abi.Billboard * max(n, 1))(*billboards)
        out = (abi.BillboardPacked * max(n, 1))()
        got = C.c_uint32(0)
        _check(lib.sprBillboardPack(arr, n, out, C.byref(got)))
        return out, got.value

    def add_packed(self, pointer, count, device=False):
        """sprBillboardAddPacked: `pointer` = address of `count` SprBillboardPacked in host (device=False) or device memory;
        the caller keeps the array alive until render() returns."""
        _check(lib.sprBillboardAddPacked(self.h, C.c_void_p(pointer), count, 1 if device else 0))

    def render(self, render_args):
        """renderMain: returns (vertices [4 * quads] structured array copied from the device, draws [(atlas, count, firstQuad)],
        worldToClip)."""
        out = abi.BillboardOutput()
        _check(lib.sprBillboardRenderMain(self.h, C.byref(render_args), C.byref(out)))
        verts = np.zeros(4 * out.numQuads, dtype=abi.BILLBOARD_VERTEX_DTYPE)
        got = C.c_uint32(0)
        _check(lib.sprBillboardReadVertices(self.h, verts.ctypes.data, out.numQuads, C.byref(got)))
        assert got.value == out.numQuads
        draws = [(out.draws[i].atlas, out.draws[i].count, out.draws[i].firstQuad) for i in range(out.numDraws)]
        return verts, draws, list(out.worldToClip)


class ParticleSystem:
    """cl::ParticleSystem over the C ABI (same calls as oracle.orc.OracleSystem)."""

    def __init__(self, ctx, seed=(1, 2, 3, 4)):
        self.ctx = ctx
        desc = abi.make_systems_desc(seed)
        self.h = C.c_void_p()
        _check(lib.sprSystemCreate(ctx.h, C.byref(desc), C.byref(self.h)))
        ctx.systems.append(self)
        self._keep = []

    def close(self):
        if self.h:
            lib.sprSystemDestroy(self.h)
            self.h = None
            if self in self.ctx.systems:
                self.ctx.systems.remove(self)

    def reserve_type(self, comp):
        self._keep.append(comp)
        t = C.c_uint32()
        _check(lib.sprReserveEffectType(self.h, C.byref(comp), C.byref(t)))
        return t.value

    def release_type(self, comp):
        _check(lib.sprReleaseEffectType(self.h, C.byref(comp)))

    def add_effect(self, comp, transform=None, entity_id=0, component_index=0):
        self._keep.append(comp)
        t = transform if transform is not None else abi.make_transform()
        e = C.c_uint32()
        _check(lib.sprAddEffect(self.h, entity_id, component_index, C.byref(comp), C.byref(t), C.byref(e)))
        return e.value

    def update_transform(self, effect_id, update):
        _check(lib.sprUpdateTransform(self.h, effect_id, C.byref(update)))

    def prepare_for_remove(self, effect_id, frame_args):
        r = C.c_int()
        _check(lib.sprPrepareForRemove(self.h, effect_id, C.byref(frame_args), C.byref(r)))
        return bool(r.value)

    def remove(self, effect_id):
        _check(lib.sprRemoveEffect(self.h, effect_id))

    def update(self, ids, frame_args):
        arr, p = _ids(ids)
        _check(lib.sprUpdateParticles(self.h, p, len(arr), C.byref(frame_args)))

    def render(self, ids, render_args, capacity=None):
        arr, p = _ids(ids)
        out = abi.RenderOutput()
        _check(lib.sprRenderMain(self.h, p, len(arr), C.byref(render_args), C.byref(out)))
        recs = []
        for i in range(out.numRecords):
            r = abi.DrawRecord()
            C.memmove(C.byref(r), C.byref(out.records[i]), C.sizeof(abi.DrawRecord))
            recs.append(r)
        self.last_device_vertices = out.deviceVertices
        return recs, out.numSortedEffects

    def query_effects(self, frustum, capacity=None):
        """Ids of the effects whose sphere area intersects `frustum`, ascending (sprQueryEffects =
        AreaSystem::queryFrustum over the particle spheres, src/client/AreaSystem.cpp:139-170)."""
        count = C.c_uint32(0)
        if capacity is None:
            _check(lib.sprQueryEffects(self.h, C.byref(frustum), None, 0, C.byref(count)))
            capacity = count.value
        out = np.zeros(max(capacity, 1), dtype=np.uint32)
        _check(lib.sprQueryEffects(self.h, C.byref(frustum), out.ctypes.data_as(C.POINTER(C.c_uint32)), capacity, C.byref(count)))
        return out[:min(count.value, capacity)].copy(), count.value

    def shade_vertices(self, records, device_out, capacity_vertices, fog=None):
        """Particle.glsl `vs` over draw records (sprShadeVertices); device_out is a device pointer with room for
        capacity_vertices SprShadedVertex. Returns the number of vertices written (stream-ordered, not synchronised)."""
        n = len(records)
        arr = (abi.DrawRecord * max(n, 1))(*records)
        total = C.c_uint64(0)
        _check(lib.sprShadeVertices(self.h, arr, n, C.byref(fog) if fog is not None else None, device_out, capacity_vertices, C.byref(total)))
        return total.value

    def effect_info(self, effect_id):
        info = abi.EffectInfo()
        _check(lib.sprGetEffectInfo(self.h, effect_id, C.byref(info)))
        return info

    def _read(self, fn, effect_id, dtype):
        n = C.c_uint32()
        _check(fn(self.h, effect_id, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=dtype)
        _check(fn(self.h, effect_id, out.ctypes.data, n.value, C.byref(n)))
        return out[:n.value]

    def read_free(self, effect_id):
        return self._read(lib.sprReadFreeIndices, effect_id, np.uint32)

    def read_stream(self, effect_id):
        return self._read(lib.sprReadVertexStream, effect_id, GPU_PARTICLE_DTYPE)

    def read_slots(self, effect_id):
        return self._read(lib.sprReadSlots, effect_id, GPU_PARTICLE_DTYPE)

    def read_stream_slots(self, effect_id):
        return self._read(lib.sprReadStreamSlots, effect_id, np.uint32)

    def rng_state(self):
        out = (C.c_uint64 * 2)()
        _check(lib.sprGetRngState(self.h, out))
        return int(out[0]), int(out[1])

    def type_lut(self, type_id):
        out = np.zeros(abi.SPLINE_LUT_BYTES, dtype=np.uint8)
        _check(lib.sprGetEffectTypeLut(self.h, type_id, out.ctypes.data))
        return out

    def type_ubo(self, type_id):
        out = abi.VertexTypeUbo()
        _check(lib.sprGetEffectTypeUbo(self.h, type_id, C.byref(out)))
        return out

    def device_view(self, effect_id):
        v = abi.EffectDeviceView()
        _check(lib.sprGetEffectDeviceView(self.h, effect_id, C.byref(v)))
        return v
