#!/usr/bin/env python3
"""bench.py -- particle-steps/s of the update + sort + billboard step (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4]

One "step" = one pass of the hot path over every particle of the workload:
sprUpdateParticles[Batch] = plan + emit + integrate/compact/depth-key + segmented radix
sort + 4x vertex expansion (sort mode on), through the C ABI of include/spear_particles.h.

Workloads:
  c3  (default at N=1)  BASELINE configs[2]: one effect, 10M live particles -- the configuration
                        north_star quotes the 70 %-of-HBM-roofline target on.
  c2                    BASELINE configs[1]: 64 emitters x 16k particles, 4 effect types.
  c4  (default at N>1)  BASELINE configs[3]: 1024 independent instances (64M particles) at 8 GPUs =
                        128 instances x 65536 particles per GPU (weak scaling: per-GPU work fixed).
                        configs[3] is defined by the all-gather of the sorted runs into one draw stream, so at
                        N > 1 `value` = the rate WITH that assembly on every GPU (fused peer-memory expansion,
                        overlapped with the next step, checked bit for bit against the NCCL path); the shard
                        compute alone is reported beside it under "compute_only", every assembly variant
                        under "gather". At N = 1 the same per-GPU workload is measured as "c4_slice".

--impl reference times the reference's own CPU implementation (oracle/_ref, built from the
reference sources; falls back to the plain-C port oracle/libspear_oracle.so) on the host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CAMERA = (16.0, 20.0, -30.0)
METRIC = "particle-steps/sec (update+sort+billboard)"
UNIT = "particle-steps/s"
ALG_BYTES_SORTED = 220.0   # 60 update + 32 re-read in sorted order + 128 vertices
ALG_BYTES_UNSORTED = 188.0
# per-kernel algorithmic bytes per particle (DESIGN.md "Kernels"): pairs (key,slot) are excluded by section 8(d)
KERNEL_ALG_BYTES = {"k_integrate": 60.0, "k_expand_sorted": 160.0, "k_expand_stream": 160.0, "k_stream_fused": 188.0}


TRAFFIC_FILES = ("r02_traffic.json", "r01_traffic.json")   # newest committed capture first


def profiled_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (C3 only)."""
    for name in TRAFFIC_FILES:
        try:
            v = json.load(open(os.path.join(ROOT, "profiles", name))).get(kernel)
            if v is not None:
                return v, "profiles/%s (ncu --set full, dram bytes per launch)" % name
        except Exception:
            pass
    return None, None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------
# workload descriptors (shared by both arms)
# ---------------------------------------------------------------------------

def workload_components(name):
    from spear_b200 import abi
    import scenarios as S
    if name == "c3":
        comp = abi.make_component(
            timeStep=1 / 60, burstAmount=10_000_000, spawnTime=1e9, lifeTime=1e9, drag=0.3, gravity=(0, -9.81, 0),
            emitPosition=dict(boxExtent=(8, 4, 8)), emitVelocity=dict(sphere=dict(minRadius=1, maxRadius=3)),
            scaleSpline=[(0, 0), (0.3, 1), (1, 0.2)], gradient=dict(points=[(0, (1, 0.5, 0.1)), (1, (0.1, 0.1, 0.1))]))
        return dict(systems=1, effects=[(comp, abi.make_transform((1, 2, 3)))], particles_per_effect=10_000_000,
                    desc="C3: 1 effect x 10M particles (BASELINE configs[2]), burst-populated, steady population")
    if name == "c2":
        effects = []
        comps = [S.comp_c2(k, burst=16384) for k in range(4)]
        for i in range(64):
            effects.append((comps[i % 4], abi.make_transform((4.0 * (i % 8), 0.0, 4.0 * (i // 8)))))
        return dict(systems=1, effects=effects, particles_per_effect=16384,
                    desc="C2: 64 emitters x 16k particles, 4 effect types (BASELINE configs[1])")
    if name == "c5":
        return dict(systems=1000, effects_per_system=100, comps=[S.comp_c5()], particles_per_effect=40,
                    desc="C5 slice: 1000 systems x 100 effects, 20 die + 20 spawn per effect-step, ~40 live per effect (BASELINE configs[4] = 32M particles at 8 GPUs)")
    if name == "c4":
        comps = [S.comp_c2(k % 4, burst=65536) for k in range(8)]
        for k in range(4, 8):
            comps[k].drag = 0.35
            comps[k].gravity.y = -4.0
        return dict(systems=128, comps=comps, particles_per_effect=65536,
                    desc="C4 slice: 128 independent systems x 1 effect x 65536 particles per GPU (BASELINE configs[3] = 1024 instances / 64M at 8 GPUs)")
    raise SystemExit("unknown workload %s" % name)


# ---------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------

class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Summarise the samples taken in [t0, t1] (perf_counter), all samples if no window is given."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if t0 is not None and not (t0 <= ts <= t1 + 0.05):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------

def build_ours(ctx, wl):
    """Create systems/effects; returns (systems, id lists, dt)."""
    from spear_b200 import abi
    systems, ids = [], []
    if "effects" in wl:
        s = ctx.create_system((1, 2, 3, 4))
        my = [s.add_effect(c, t) for c, t in wl["effects"]]
        systems.append(s); ids.append(my)
    else:
        for i in range(wl["systems"]):
            s = ctx.create_system((1, 2 + i + wl.get("seed_offset", 0), 3, 4))
            comp = wl["comps"][i % len(wl["comps"])]
            my = [s.add_effect(comp, abi.make_transform((4.0 * (i % 16) + 0.25 * k, 0.0, 4.0 * (i // 16))))
                  for k in range(wl.get("effects_per_system", 1))]
            systems.append(s); ids.append(my)
    # dt = the largest per-effect timeStep (effects are jittered +0..10 %, ParticleSystem.cpp:744):
    # every effect does >= 1 sub-step per frame, a few do 2 now and then; particle-steps are counted exactly.
    dt = max(s.effect_info(my[0]).timeStep for s, my in zip(systems, ids)) * (1.1 if wl.get("effects_per_system", 1) > 1 else 1.0)
    if "effects" in wl:
        dt = max(s.effect_info(e).timeStep for s, my in zip(systems, ids) for e in my)
    return systems, ids, dt


def run_ours(args, rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from spear_b200 import abi, particles

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    json_fd = None
    if world > 1:
        # NCCL prints its version line to the process's stdout (whatever NCCL_DEBUG_FILE says): from here on file
        # descriptor 1 is stderr for every library in the process, and the one JSON line goes to the saved stdout
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name = args.workload or ("c3" if world == 1 else "c4")
    wl = workload_components(name)
    if name in ("c4", "c5"):
        wl["seed_offset"] = rank * wl["systems"]
    total_particles = wl["particles_per_effect"] * (len(wl["effects"]) if "effects" in wl else wl["systems"] * wl.get("effects_per_system", 1))
    if name == "c5":
        total_particles = wl["systems"] * wl["effects_per_system"] * 128
    ctx = particles.ParticleContext(device=local, sort_particles=not args.no_sort, max_particles_per_frame=1 << 30,
                                    initial_slot_capacity=int(total_particles * 1.1) + (1 << 16))
    ext = torch.cuda.ExternalStream(ctx.stream, device=local)
    systems, ids, dt = build_ours(ctx, wl)
    batch = ctx.make_batch(systems, ids)
    all_effects = [(s, e) for s, my in zip(systems, ids) for e in my]

    frame = [0]

    def step():
        fa = abi.make_frame_args(dt, game_time=frame[0] * dt, frame_index=frame[0], camera=CAMERA)
        ctx.update_prepared(batch, fa)
        frame[0] += 1

    def sim_steps():
        return ctx.totals(alive=False)[0]

    def alive():
        return ctx.totals()[1]

    sampler = ClockSampler(local)
    sampler.start()
    # population: burst step(s), untimed
    for _ in range(6 if name == "c5" else 2):
        step()
    ctx.synchronize()
    for _ in range(args.warmup):
        step()
    ctx.synchronize()
    n_alive = alive()

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- timed regions: exactly K steps each, CUDA events on the launching stream, barrier + synchronize on both
    # sides. The region is repeated `--repeats` times back to back (the workload is steady: same population, same
    # launches) and `value` is the MEDIAN region; every region's time is listed under "repeats".
    regions = []
    tw0 = time.perf_counter()
    for _ in range(max(1, args.repeats)):
        steps0, launches0 = sim_steps(), ctx.launch_count
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(ext)
        for _ in range(args.steps):
            step()
        ev1.record(ext)
        barrier()
        r_ms = ev0.elapsed_time(ev1)
        if world > 1:   # the region's time is the slowest rank's
            t_r = torch.tensor([r_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t_r, op=dist.ReduceOp.MAX)
            r_ms = float(t_r.item())
        regions.append((r_ms, ctx.launch_count - launches0, sim_steps() - steps0))
    tw1 = time.perf_counter()
    ms, launches, effect_steps = sorted(regions)[len(regions) // 2]
    repeats = {"n": len(regions), "ms_per_step": [r[0] / args.steps for r in regions], "median_ms_per_step": ms / args.steps,
               "min_ms_per_step": min(r[0] for r in regions) / args.steps, "max_ms_per_step": max(r[0] for r in regions) / args.steps,
               "note": "value and ms_per_step are the median of these regions of exactly %d steps each" % args.steps}
    # per-kernel durations: the same K steps again, immediately, with a CUDA event pair around every launch
    # (the events cost 5-40 % of the step on these short kernels, so they are kept out of `value`'s region)
    ktimes = {}
    if not args.no_kernel_timing:
        ctx.enable_kernel_timing(True)
        ctx.kernel_times(reset=True)
        ks0 = sim_steps()
        for _ in range(args.steps):
            step()
        ctx.synchronize()
        ktimes = ctx.kernel_times(reset=True)
        ctx.enable_kernel_timing(False)
        kernel_region_particle_steps = (sim_steps() - ks0) * (n_alive / len(all_effects))
    # nvidia-smi samples every 100 ms; a timed region shorter than that gets its clocks from the same
    # steps repeated back to back right after it (identical load, not part of any reported time)
    clock_window = (tw0, tw1)
    if tw1 - tw0 < 0.7:
        tx0 = time.perf_counter()
        while time.perf_counter() - tx0 < 0.8:
            for _ in range(args.steps):
                step()
            ctx.synchronize()
        clock_window = (tw0, time.perf_counter())
        steps_after_clock_loop = True
    else:
        steps_after_clock_loop = False
    # every effect holds a constant population in the timed window; sub-steps are counted exactly
    per_effect_alive = n_alive / len(all_effects)
    particle_steps = effect_steps * per_effect_alive

    # ---- e2e: the plugin call a user makes, host buffers in, host-visible result out, every step:
    # sprUpdateParticlesBatch (H2D of the frame's work lists from pinned memory) + sprRenderMain
    # (D2H of the per-effect counts/bounds, host effect sort, draw records) + synchronise.
    # Multi-system workloads use the batched forms (one read-back for all systems).
    from tests_render_args import make_render_args  # noqa: E402
    ra = make_render_args()
    for k in range(4):  # an all-inclusive frustum: the bench population falls for as long as the bench runs
        for pl in range(4):
            ra.frustum.side[k][pl] = 1.0 if k == 3 else 0.0
            ra.frustum.caps[k][pl] = 1.0 if k == 3 else 0.0
    vis = [np.ascontiguousarray(my, dtype=np.uint32) for my in ids]
    barrier()
    e_steps0 = sim_steps()
    t0 = time.perf_counter()
    n_records = 0
    for _ in range(args.steps):
        step()
        outs = ctx.render_prepared(batch, ra)
        n_records += sum(o.numRecords for o in outs)
    ctx.synchronize()
    assert n_records > 0, "renderMain produced no draw records"
    t1 = time.perf_counter()
    clocks = sampler.stop(*clock_window)
    clocks["window"] = "timed region" + (" + the same steps repeated for 0.8 s right after it (region shorter than the 100 ms sampling period)" if steps_after_clock_loop else "")
    e2e_ms = (t1 - t0) * 1e3
    e2e_particle_steps = (sim_steps() - e_steps0) * per_effect_alive
    n_eff = len(all_effects)
    h2d = n_eff * 16 + len(systems) * 16 + n_eff * 4  # ActiveEntry + SystemWork per frame, ids for the gather
    d2h = n_eff * 60                                   # RenderGather per visible effect

    # ---- e2e, strictest reading: the vertex stream itself is copied to (pinned) host memory every step as well,
    # i.e. the product is used like the reference's CPU path, whose output array lives in host memory before
    # sg_append_buffer uploads it. A GPU renderer never needs this copy; it is reported, not used as the headline.
    e2e_host = None
    if world == 1 and n_alive * 128 < 8e9 and len(all_effects) <= 256:
        import ctypes as C
        hostbuf = torch.empty(int(n_alive * 1.02 + 4096) * 128, dtype=torch.uint8, pin_memory=True)
        cap_vertices = hostbuf.numel() // 32
        cnt = C.c_uint32(0)
        hs0 = sim_steps()
        th0 = time.perf_counter()
        hsteps = min(args.steps, 3)
        copied = 0
        for _ in range(hsteps):
            step()
            ctx.render_prepared(batch, ra)
            off = 0
            for s_, e_ in all_effects:
                particles._check(particles.lib.sprReadVertexStream(s_.h, e_, hostbuf.data_ptr() + off * 32, cap_vertices - off, C.byref(cnt)))
                off += cnt.value
            copied += off * 32
        th1 = time.perf_counter()
        e2e_host = {"value": (sim_steps() - hs0) * per_effect_alive / (th1 - th0), "unit": UNIT, "ms_per_step": (th1 - th0) * 1e3 / hsteps,
                    "d2h_bytes_per_step": copied // hsteps,
                    "note": "as e2e, plus sprReadVertexStream of every effect into pinned host memory each step (%d steps)" % hsteps}
        del hostbuf

    # ---- vertex stage: Particle.glsl `vs` over the draw records of the last frame, N=1 only.
    # 32 B read + 4 x 64 B written per particle; timed alone (it is not part of `value`).
    vertex_stage = None
    if world == 1 and not args.no_vertex_stage and len(outs) == 1 and n_alive * 256 < 40e9:
        recs = [outs[0].records[i] for i in range(outs[0].numRecords)]
        nv = 4 * sum(r.numParticles for r in recs)
        shaded = torch.empty((nv, 16), dtype=torch.float32, device="cuda")
        for _ in range(3):
            systems[0].shade_vertices(recs, shaded.data_ptr(), nv)
        ctx.synchronize()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        v0.record(ext)
        for _ in range(reps):
            systems[0].shade_vertices(recs, shaded.data_ptr(), nv)
        v1.record(ext)
        ctx.synchronize(); torch.cuda.synchronize()
        vms = v0.elapsed_time(v1) / reps
        peak_v, _src = measured_peak()
        vertex_stage = {"kernel": "k_shade_vertices", "ms_per_launch": vms, "vertices": nv, "particles_per_s": nv / 4 / (vms * 1e-3),
                        "alg_bytes_per_particle": 288.0, "achieved_gbs": nv / 4 * 288.0 / (vms * 1e-3) / 1e9,
                        "frac": nv / 4 * 288.0 / (vms * 1e-3) / 1e9 / peak_v,
                        "note": "sprShadeVertices over the frame's draw records: gl_Position, uv0, uv1, colour, params (64 B) for 4 vertices per particle"}
        del shaded

    # ---- multi-GPU draw-stream assembly: every rank ends up with the whole draw stream (4 vertices per
    # particle, rank-major). Two implementations, both timed:
    #   nccl : sprPackRuns -> NCCL all_gather of the 32-byte records -> sprExpandRuns
    #   fused: sprRunBufferPack -> 4-byte all_reduce as the cross-rank barrier -> sprExpandFromPeers, ONE kernel
    #          that reads every rank's run out of peer memory over NVLink and expands it while it arrives
    gather = None
    if world > 1:
        cap = int(n_alive * 1.02) + 1024
        mine = torch.empty((cap, 8), dtype=torch.float32, device="cuda")
        allrec = torch.empty((world * cap, 8), dtype=torch.float32, device="cuda")
        verts = torch.empty((world * cap * 4, 8), dtype=torch.float32, device="cuda")
        sys_list = [s for s, e in all_effects]
        eff_list = [e for s, e in all_effects]
        pack_list = ctx.make_pack_list(sys_list, eff_list)
        bufs = [ctx.run_buffer_create(cap) for _ in range(2)]
        my_handles = [ctx.run_buffer_ipc_handle(b) for b in bufs]
        all_handles = [None] * world
        dist.all_gather_object(all_handles, my_handles)
        peer_ptrs = [[ctx.run_buffer_ptr(bufs[k]) if r == rank else ctx.run_buffer_open_peer(all_handles[r][k]) for r in range(world)]
                     for k in range(2)]
        flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        verts_f = torch.empty((world * cap * 4, 8), dtype=torch.float32, device="cuda")
        total_dev = torch.zeros(1, dtype=torch.int64, device="cuda")
        kstep = [0]
        with torch.cuda.stream(ext):
            def gstep():
                step()
                ctx.pack_runs(sys_list, eff_list, mine.data_ptr(), cap, None, want_total=False)
                dist.all_gather_into_tensor(allrec, mine)
                ctx.expand_runs(allrec.data_ptr(), world * cap, verts.data_ptr())

            def fstep(do_step=True):
                k = kstep[0] & 1
                kstep[0] += 1
                if do_step:
                    step()
                ctx.run_buffer_pack(pack_list, bufs[k])
                dist.all_reduce(flag)           # stream-ordered cross-rank barrier: every rank's pack is complete
                ctx.expand_from_peers(peer_ptrs[k], verts_f.data_ptr(), world * cap, total_dev.data_ptr(), my_rank=rank)

            def timed(fn):
                for _ in range(2):
                    fn()
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                gs0 = sim_steps()
                g0.record(ext)
                for _ in range(args.steps):
                    fn()
                g1.record(ext)
                barrier()
                return g0.elapsed_time(g1), (sim_steps() - gs0) * per_effect_alive

            # fused + overlapped: the assembly of frame k (barrier + peer-memory expansion) runs on a second stream
            # while the first one already simulates frame k+1; the two run buffers carry the dependency
            ctx_b = particles.ParticleContext(device=local, initial_slot_capacity=1024)   # only its stream is used
            ext_b = torch.cuda.ExternalStream(ctx_b.stream, device=local)
            ev_pack = [torch.cuda.Event() for _ in range(2)]
            ev_free = [torch.cuda.Event() for _ in range(2)]
            ostate = {"k": 0}

            def ostep():
                n = ostate["k"]
                k = n & 1
                ostate["k"] = n + 1
                if n >= 2:
                    ext.wait_event(ev_free[k])   # every rank has finished the expansion that read buffer k
                step()
                ctx.run_buffer_pack(pack_list, bufs[k])
                ev_pack[k].record(ext)
                ext_b.wait_event(ev_pack[k])
                with torch.cuda.stream(ext_b):
                    dist.all_reduce(flag)        # all packs of frame n are complete, and so are all expansions of frame n-1
                ev_free[k ^ 1].record(ext_b)
                ctx_b.expand_from_peers(peer_ptrs[k], verts_f.data_ptr(), world * cap, total_dev.data_ptr(), my_rank=rank)

            def timed_overlapped(share):
                """`--repeats` regions of exactly K steps each (simulation of frame k+1 under the assembly of frame k);
                every region ends when the LAST assembly has finished on every rank. Returns the median region.
                share = resident CTAs per SM of the assembly kernel (sprContextSetAssemblyShare): it is NVLink-bound with
                2 CTAs per SM as with 8, and what it leaves free runs the next frame's simulation."""
                ctx_b.set_assembly_share(share)
                for _ in range(3):
                    ostep()
                regs = []
                for _ in range(max(1, args.repeats)):
                    ctx_b.synchronize(); barrier()
                    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    done = torch.cuda.Event()
                    gs0, l0 = sim_steps(), ctx.launch_count + ctx_b.launch_count
                    g0.record(ext)
                    ext_b.wait_event(g0)
                    for _ in range(args.steps):
                        ostep()
                    done.record(ext_b)
                    ext.wait_event(done)
                    g1.record(ext)
                    ctx_b.synchronize(); barrier()
                    r_ms = g0.elapsed_time(g1)
                    t_r = torch.tensor([r_ms], dtype=torch.float64, device="cuda")   # the region's time is the slowest rank's
                    dist.all_reduce(t_r, op=dist.ReduceOp.MAX)
                    regs.append((float(t_r.item()), (sim_steps() - gs0) * per_effect_alive, ctx.launch_count + ctx_b.launch_count - l0))
                return sorted(regs)[len(regs) // 2], [r[0] / args.steps for r in regs]

            gms, gps = timed(gstep)
            fms, fps = timed(fstep)
            by_share = {}
            for share in [int(x) for x in args.assembly_shares.split(",")]:
                by_share[share] = timed_overlapped(share)
            best_share = min(by_share, key=lambda k: by_share[k][0][0])
            (oms, ops, olaunches), o_regions = by_share[best_share]
            ctx_b.synchronize()
            # the assembly kernel alone, all ranks pulling at the same time (a 4-byte all_reduce lines them up): the
            # NVLink roofline of sprExpandFromPeers
            ak = []
            for _ in range(args.steps):
                dist.all_reduce(flag)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(ext)
                ctx.expand_from_peers(peer_ptrs[(kstep[0] - 1) & 1], verts_f.data_ptr(), world * cap, total_dev.data_ptr(), my_rank=rank)
                a1.record(ext)
                ak.append((a0, a1))
            barrier()
            assembly_kernel_ms = sum(a.elapsed_time(b) for a, b in ak) / len(ak)
            # same state through both paths: the fused stream must equal the NCCL one (padding removed)
            ctx.pack_runs(sys_list, eff_list, mine.data_ptr(), cap, None, want_total=False)
            dist.all_gather_into_tensor(allrec, mine)
            ctx.expand_runs(allrec.data_ptr(), world * cap, verts.data_ptr())
            fstep(do_step=False)
            barrier()
            counts = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
            dist.all_gather(counts, torch.tensor([int(n_alive)], dtype=torch.int64, device="cuda"))
            counts = [int(c.item()) for c in counts]
            ok, off = True, 0
            for r, c in enumerate(counts):
                a = verts_f[off * 4:(off + c) * 4].view(torch.int32)
                b = verts[r * cap * 4:(r * cap + c) * 4].view(torch.int32)
                ok = ok and bool(torch.equal(a, b))
                off += c
            ok = ok and int(total_dev.item()) == sum(counts)
            # deferred expansion (sprContextSetDeferredExpansion): the consumer is the assembled stream, which holds this
            # GPU's shard too, so the updates sort but do not write the local 4x stream and the pack gathers the sorted
            # records from the slot state. Same overlapped pipeline, best share; then the same state through the NCCL pair
            # and the fused kernel again (both now packing from the slot state) must give the same stream. That the pack
            # from the slot state equals the pack from the local stream is tests/test_gpu_more.py::test_deferred_expansion_*.
            ctx.set_deferred_expansion(True)
            d_by_share = {}
            # both assembly kernels (sprContextSetAssemblyKernel: 1 = load/store, 2 = bulk-asynchronous pipeline) at every share: which
            # one wins depends on what bounds the assembly at this GPU count (NVLink ingress at 8, HBM stores at 2-4)
            for kern in (1, 2):
                ctx_b.set_assembly_kernel(kern)
                for share in [int(x) for x in args.assembly_shares.split(",")]:   # the lighter simulation may want a different share
                    d_by_share[(share, kern)] = timed_overlapped(share)
            d_best = min(d_by_share, key=lambda k: d_by_share[k][0][0])
            d_best_share = d_best[0]
            (dms, dps, dlaunches), d_regions = d_by_share[d_best]
            ctx_b.set_assembly_kernel(d_best[1])
            ctx_b.synchronize()
            ctx.pack_runs(sys_list, eff_list, mine.data_ptr(), cap, None, want_total=False)
            dist.all_gather_into_tensor(allrec, mine)
            ctx.expand_runs(allrec.data_ptr(), world * cap, verts.data_ptr())
            ctx.set_assembly_kernel(d_best[1])   # the check runs the kernel that was timed
            verts_f.zero_()
            fstep(do_step=False)
            ctx.set_assembly_kernel(0)
            barrier()
            ok_d, off = True, 0
            for r, c in enumerate(counts):
                a = verts_f[off * 4:(off + c) * 4].view(torch.int32)
                b = verts[r * cap * 4:(r * cap + c) * 4].view(torch.int32)
                ok_d = ok_d and bool(torch.equal(a, b))
                off += c
            ok_d = ok_d and int(total_dev.item()) == sum(counts)
            ctx.set_deferred_expansion(False)
            step()   # the local streams are current again
            ctx.synchronize()
        gather = {"deferred_ms_per_step": dms / args.steps, "deferred_particle_steps_this_rank": dps, "deferred_launches_this_rank": dlaunches,
                  "deferred_regions_ms_per_step": d_regions, "deferred_matches_nccl": ok_d,
                  "deferred_share": d_best_share, "deferred_kernel": "bulk_async" if d_best[1] == 2 else "load_store",
                  "deferred_by_share": {"%d_%s" % (k[0], "bulk_async" if k[1] == 2 else "load_store"): v[0][0] / args.steps for k, v in d_by_share.items()},
                  "ms_per_step": gms / args.steps, "particle_steps_this_rank": gps, "fused_ms_per_step": fms / args.steps,
                  "fused_particle_steps_this_rank": fps, "fused_matches_nccl": ok,
                  "overlapped_ms_per_step": oms / args.steps, "overlapped_particle_steps_this_rank": ops,
                  "overlapped_launches_this_rank": olaunches, "overlapped_regions_ms_per_step": o_regions,
                  "overlapped_share": best_share, "overlapped_by_share": {str(k): v[0][0] / args.steps for k, v in by_share.items()},
                  "assembly_kernel_ms": assembly_kernel_ms,
                  "nvlink_bytes_in_per_step": int(sum(c for r, c in enumerate(counts) if r != rank) * 32)}

    # ---- max over ranks / aggregate
    if world > 1:
        t = torch.tensor([ms, e2e_ms, gather["ms_per_step"] * args.steps, gather["fused_ms_per_step"] * args.steps, gather["assembly_kernel_ms"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, gms_tot, fms_tot, akms = [float(x) for x in t.tolist()]
        c = torch.tensor([particle_steps, e2e_particle_steps, gather["particle_steps_this_rank"], float(launches),
                          gather["fused_particle_steps_this_rank"], 1.0 if gather["fused_matches_nccl"] else 0.0,
                          gather["overlapped_particle_steps_this_rank"], float(gather["overlapped_launches_this_rank"]),
                          gather["deferred_particle_steps_this_rank"], float(gather["deferred_launches_this_rank"]),
                          1.0 if gather["deferred_matches_nccl"] else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        particle_steps, e2e_particle_steps, g_total_ps, launches_all, f_total_ps, f_ok, o_total_ps, o_launches_all, d_total_ps, d_launches_all, d_ok = [float(x) for x in c.tolist()]
        dms = gather["deferred_ms_per_step"] * args.steps   # already the max over ranks (timed_overlapped reduces every region)
    else:
        launches_all = float(launches)

    if rank == 0:
        peak, peak_src = measured_peak()
        value = particle_steps / (ms * 1e-3)
        sorted_mode = not args.no_sort
        alg = ALG_BYTES_SORTED if sorted_mode else ALG_BYTES_UNSORTED
        # dominant kernel by measured time
        kern = {k: {"ms_total": v[0], "launches": v[1]} for k, v in ktimes.items()}
        tot_k = sum(v["ms_total"] for v in kern.values()) or 1.0
        alg_key = dict(KERNEL_ALG_BYTES)
        for k, v in kern.items():
            v["share"] = v["ms_total"] / tot_k
            v["ms_per_launch"] = v["ms_total"] / max(v["launches"], 1)
            if k in alg_key:
                # particles processed by one launch of this kernel (this rank)
                per_launch = kernel_region_particle_steps / max(v["launches"], 1)
                v["alg_bytes_per_launch"] = alg_key[k] * per_launch
                v["achieved_gbs"] = v["alg_bytes_per_launch"] / (v["ms_per_launch"] * 1e-3) / 1e9
        cand = [k for k in kern if "achieved_gbs" in kern[k]]
        dom = max(cand, key=lambda k: kern[k]["ms_total"]) if cand else None
        roof = None
        if dom:
            roof = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kern[dom]["achieved_gbs"] / peak, "traffic": profiled_traffic(dom)[0] if name == "c3" else None,
                    "traffic_source": profiled_traffic(dom)[1] if name == "c3" else None,
                    "peak_source": peak_src,
                    "alg_bytes_per_particle": alg_key[dom],
                    "step": {"alg_bytes_per_particle_step": alg, "achieved": value / world * alg / 1e9,
                             "frac": value / world * alg / 1e9 / peak,
                             "note": "whole step per GPU: particle-steps/s x %g B (SURVEY.md 8(d)) / peak" % alg}}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "name": name, "particles_per_gpu": int(n_alive), "sort": sorted_mode,
                       "l2": "inputs larger than L2 (state %d MB + vertex stream %d MB per GPU)" % (n_alive * 32 >> 20, n_alive * 128 >> 20)
                       if n_alive * 160 > 200e6 else "working set %d MB: state is L2-resident, the 128 B/particle vertex stream is not re-read" % (n_alive * 160 >> 20),
                       "dt": dt, "parallelism": "independent systems sharded per GPU" if world > 1 else "1 GPU"},
            "e2e": {"value": e2e_particle_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps,
                    "note": "sprUpdateParticlesBatch + sprRenderMainBatch per step through the C ABI with host id lists in and host draw records out; "
                            "the vertex stream stays in HBM where the renderer reads it (the reference uploads it with sg_append_buffer)",
                    "with_vertex_stream_to_host": e2e_host},
            "gpu_launches": int(launches_all),
            "repeats": repeats,
            "clocks": clocks,
            "roofline": roof,
            "kernels": kern,
            "kernel_timing": "CUDA event pair around every launch, over the same K steps repeated immediately after the timed region",
        }
        if vertex_stage is not None:
            out["vertex_stage"] = vertex_stage
        if gather is not None:
            # N > 1: BASELINE configs[3] is DEFINED by the all-gather of the sorted runs into one draw stream, so the
            # headline `value` is the rate WITH the assembly (fused + overlapped path, bit-identical to the NCCL path,
            # checked below); the simulation alone -- independent shards, no communication -- is kept beside it.
            out["compute_only"] = {"value": value, "unit": UNIT, "ms_per_step": ms / args.steps, "repeats": repeats, "gpu_launches": int(launches_all),
                                   "note": "emit + update + compact + sort + local 4x expansion of each GPU's own shard, no draw-stream assembly"}
            if f_ok == world:
                out["value"] = o_total_ps / (oms * 1e-3)
                out["ms_per_step"] = oms / args.steps
                out["gpu_launches"] = int(o_launches_all)
                out["repeats"] = {"n": len(gather["overlapped_regions_ms_per_step"]), "ms_per_step": gather["overlapped_regions_ms_per_step"],
                                  "median_ms_per_step": oms / args.steps,
                                  "note": "value and ms_per_step are the median of these regions of exactly %d steps each, max over ranks per region" % args.steps}
                out["config"]["assembly"] = ("every GPU assembles the whole %d-GPU draw stream each step: sprRunBufferPack -> barrier -> sprExpandFromPeers "
                                             "(peer-memory reads over NVLink fused with the 4x expansion), overlapped with the next step's simulation" % world)
            else:
                out["config"]["assembly"] = "fused path did NOT match the NCCL path: value is compute only"
            if f_ok == world and d_ok == world and dms < oms:
                # the same pipeline without the redundant local stream (the assembled stream holds the local shard too)
                out["value"] = d_total_ps / (dms * 1e-3)
                out["ms_per_step"] = dms / args.steps
                out["gpu_launches"] = int(d_launches_all)
                out["repeats"] = {"n": len(gather["deferred_regions_ms_per_step"]), "ms_per_step": gather["deferred_regions_ms_per_step"],
                                  "median_ms_per_step": dms / args.steps,
                                  "note": "value and ms_per_step are the median of these regions of exactly %d steps each, max over ranks per region" % args.steps}
                out["config"]["assembly"] += ("; deferred expansion: the updates sort but do not write the local 4x stream (the assembled stream contains this GPU's shard), "
                                              "sprRunBufferPack gathers the sorted records from the slot state")
            nv = gather["nvlink_bytes_in_per_step"]
            out["gather"] = {
                "unit": UNIT,
                "nccl": {"value_with_gather": g_total_ps / (gms_tot * 1e-3), "ms_per_step": gms_tot / args.steps,
                         "note": "step + sprPackRuns (32 B sorted records) + NCCL all_gather + sprExpandRuns (4x) of the whole %d-GPU draw stream on every GPU" % world},
                "fused": {"value_with_gather": f_total_ps / (fms_tot * 1e-3), "ms_per_step": fms_tot / args.steps,
                          "assembly_ms": (fms_tot - ms) / args.steps,
                          "nvlink_gbs_in_per_gpu_during_assembly": nv / max((fms_tot - ms) / args.steps * 1e-3, 1e-9) / 1e9,
                          "nvlink_roofline": {"kernel": "k_expand_from_peers", "ms_per_launch": akms, "achieved_gbs": nv / (akms * 1e-3) / 1e9,
                                              "peak_gbs": 770.0, "peak_source": "B200_PROFILING.md measured peer copy per direction",
                                              "frac": nv / (akms * 1e-3) / 1e9 / 770.0,
                                              "note": "the assembly kernel alone (CUDA events, max over ranks, all ranks pulling at once); all-pull-all with no stores at all "
                                                      "tops out at ~655 GB/s on this box (profiles/r02_peer_allgather_microbench_8gpu.txt)"},
                          "matches_nccl_path_bit_for_bit": f_ok == world,
                          "note": "step + sprRunBufferPack + 4-byte all_reduce barrier + sprExpandFromPeers: ONE kernel per rank reads every rank's run from IPC-mapped peer memory over NVLink and writes the 4x expanded draw stream locally"},
                "fused_overlapped": {"value_with_gather": o_total_ps / (oms * 1e-3), "ms_per_step": oms / args.steps,
                                     "nvlink_gbs_in_per_gpu": nv / (oms / args.steps * 1e-3) / 1e9,
                                     "nvlink_roofline_frac": nv / (oms / args.steps * 1e-3) / 1e9 / 770.0,
                                     "assembly_ctas_per_sm": gather["overlapped_share"], "ms_per_step_by_assembly_share": gather["overlapped_by_share"],
                                     "note": "the fused assembly of frame k runs on a second stream while frame k+1 is simulated (double-buffered run buffers); "
                                             "the assembly kernel keeps a fixed share of every SM (sprContextSetAssemblyShare), the best share of --assembly-shares is reported; "
                                             "at shares of 1 and 2 CTAs per SM it is the bulk-asynchronous pipeline (cp.async.bulk loads + tensor stores), above that the load/store kernel"},
                "fused_deferred": {"value_with_gather": d_total_ps / (dms * 1e-3), "ms_per_step": dms / args.steps,
                                   "assembly_ctas_per_sm": gather["deferred_share"], "assembly_kernel": gather["deferred_kernel"],
                                   "ms_per_step_by_assembly_share": gather["deferred_by_share"],
                                   "matches_nccl_path_bit_for_bit": d_ok == world,
                                   "nvlink_gbs_in_per_gpu": nv / (dms / args.steps * 1e-3) / 1e9,
                                   "nvlink_roofline_frac": nv / (dms / args.steps * 1e-3) / 1e9 / 770.0,
                                   "note": "fused_overlapped with sprContextSetDeferredExpansion: update + compaction + sort per shard, NO local 4x expansion "
                                           "(every GPU expands the whole assembled stream, its own shard included), the pack gathers the sorted records from the slot state"},
                "nvlink_bytes_in_per_gpu_per_step": nv}
        if world == 1 and sorted_mode and not args.no_stream_mode:
            # the strict drop-in stream (ascending slot order, no per-particle sort) on the same workload,
            # measured by the same script in a fresh process right after this one releases the GPU
            out["stream_mode"] = "pending"
        if world == 1 and sorted_mode and name == "c3" and not args.no_c4_slice:
            out["c4_slice"] = "pending"   # the per-GPU workload of the N > 1 runs on this one GPU, same script, fresh process
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(name, seconds=15.0)
    billboards = None
    if world == 1 and not args.no_billboards:
        try:
            billboards = bench_billboards(ctx, cpu_reference=not args.no_cpu_baseline)
        except Exception as ex:  # noqa: BLE001
            billboards = {"error": str(ex)[:300]}
    if world > 1:
        ctx_b.close()
        dist.destroy_process_group()
    ctx.close()
    if rank == 0:
        if billboards is not None:
            out["billboards"] = billboards
        if out.get("stream_mode") == "pending":
            try:
                cmd = [sys.executable, os.path.abspath(__file__), "--workload", name, "--no-sort", "--no-cpu-baseline", "--no-stream-mode", "--no-vertex-stage", "--no-billboards",
                       "--steps", str(args.steps), "--warmup", str(args.warmup)]
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
                sm = json.loads(r.stdout.strip().splitlines()[-1])
                out["stream_mode"] = {"value": sm["value"], "unit": UNIT, "ms_per_step": sm["ms_per_step"], "e2e": sm["e2e"]["value"],
                                      "roofline": sm["roofline"], "kernels": sm["kernels"], "gpu_launches": sm["gpu_launches"],
                                      "note": "same workload with SPR_CTX_SORT_PARTICLES off: update + compact + 4x billboard expansion in ascending slot order (the reference's own stream order)"}
            except Exception as ex:  # noqa: BLE001
                out["stream_mode"] = {"error": str(ex)[:200]}
        if out.get("c4_slice") == "pending":
            try:
                cmd = [sys.executable, os.path.abspath(__file__), "--workload", "c4", "--no-cpu-baseline", "--no-stream-mode", "--no-vertex-stage", "--no-billboards",
                       "--no-kernel-timing", "--steps", str(args.steps), "--warmup", str(args.warmup), "--repeats", str(args.repeats)]
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
                c4 = json.loads(r.stdout.strip().splitlines()[-1])
                out["c4_slice"] = {"value": c4["value"], "unit": UNIT, "ms_per_step": c4["ms_per_step"], "e2e": c4["e2e"]["value"], "repeats": c4["repeats"],
                                   "config": c4["config"], "gpu_launches": c4["gpu_launches"],
                                   "note": "the per-GPU workload of the N > 1 runs (128 systems x 65536 particles, BASELINE configs[3] slice) on this one GPU: "
                                           "the like-for-like single-GPU anchor of the scaling curve (N = 1 needs no assembly: the draw stream is already in one place)"}
            except Exception as ex:  # noqa: BLE001
                out["c4_slice"] = {"error": str(ex)[:200]}
        if json_fd is None:
            print(json.dumps(out))
        else:
            sys.stdout.flush()
            os.write(json_fd, (json.dumps(out) + "\n").encode())


# ---------------------------------------------------------------------------
# billboards: one frame of cl::BillboardSystem through the C ABI, N=1 only
# ---------------------------------------------------------------------------

def bench_billboards(ctx, n=1 << 20, frames=5, cpu_reference=True):
    """addBillboard x n + renderMain per frame with host buffers in and the host draw list out (so the H2D copy of
    the frame's billboards and the D2H copy of the atlas list are inside the time), budget = n. The reference's own
    BillboardSystem.cpp (oracle/_ref) is timed beside it on the same frame when that library is present."""
    import ctypes as C
    import numpy as np
    from spear_b200 import abi, particles
    from tests_render_args import make_render_args
    rng = np.random.default_rng(3)
    arr = np.zeros(n, dtype=np.dtype(abi.Billboard))
    arr["sprite"]["atlas"] = rng.integers(1, 9, size=n)
    arr["sprite"]["maxVert"]["x"] = 1.0; arr["sprite"]["maxVert"]["y"] = 1.0
    arr["sprite"]["vertUvScale"]["x"] = 0.25; arr["sprite"]["vertUvScale"]["y"] = 0.25
    tr = np.zeros((n, 12), dtype=np.float32)
    tr[:, 0] = tr[:, 4] = tr[:, 8] = 1.0
    tr[:, 9:12] = rng.uniform(-50, 50, size=(n, 3))
    arr["transform"] = tr
    for k in "xyzw":
        arr["color"][k] = 1.0
    arr["depth"] = rng.uniform(0, 100, size=n).astype(np.float32)
    for f in ("anchor",):
        arr[f]["x"] = 0.5; arr[f]["y"] = 0.5
    arr["cropMax"]["x"] = 1.0; arr["cropMax"]["y"] = 1.0
    ptr = arr.ctypes.data_as(C.POINTER(abi.Billboard))
    ra = make_render_args()
    bs = particles.BillboardSystem(ctx, max_billboards_per_frame=n)
    out = abi.BillboardOutput()
    times = []
    launches0 = ctx.launch_count
    for _ in range(frames + 2):
        t0 = time.perf_counter()
        particles._check(particles.lib.sprBillboardAdd(bs.h, ptr, n))
        particles._check(particles.lib.sprBillboardRenderMain(bs.h, C.byref(ra), C.byref(out)))
        times.append(time.perf_counter() - t0)
    quads, draws = out.numQuads, out.numDraws
    launches = (ctx.launch_count - launches0) // (frames + 2)
    ms = sorted(times[2:])[len(times[2:]) // 2] * 1e3
    # the same frame as packed billboards (sprBillboardPack once, outside the time): from pinned host memory (one DMA, no
    # per-billboard host work) and from device memory (a producer on the GPU / billboards that do not change: read in place)
    packed_ms = {}
    try:
        import torch
        pk = (abi.BillboardPacked * n)()
        got = C.c_uint32(0)
        particles._check(particles.lib.sprBillboardPack(ptr, n, pk, C.byref(got)))
        host = torch.from_numpy(np.frombuffer(pk, dtype=np.uint8, count=got.value * 120).copy()).pin_memory()
        dev = host.cuda()
        torch.cuda.synchronize()
        for kind, address, on_device in (("pinned_host", host.data_ptr(), 0), ("device", dev.data_ptr(), 1)):
            tt = []
            for _ in range(frames + 2):
                t0 = time.perf_counter()
                particles._check(particles.lib.sprBillboardAddPacked(bs.h, C.c_void_p(address), got.value, on_device))
                particles._check(particles.lib.sprBillboardRenderMain(bs.h, C.byref(ra), C.byref(out)))
                tt.append(time.perf_counter() - t0)
            assert out.numQuads == quads and out.numDraws == draws
            packed_ms[kind] = sorted(tt[2:])[len(tt[2:]) // 2] * 1e3
    except Exception as ex:  # noqa: BLE001
        packed_ms = {"error": str(ex)[:200]}
    bs.close()
    res = {"billboards_per_frame": n, "budget": n, "quads": int(quads), "draws": int(draws), "ms_per_frame": ms,
           "billboards_per_s": n / (ms * 1e-3), "h2d_bytes_per_frame": int(n * 120), "d2h_bytes_per_frame": int(draws * 16 + 8),
           "gpu_launches_per_frame": int(launches),
           "packed_input_ms_per_frame": packed_ms,
           "note": "sprBillboardAdd + sprBillboardRenderMain per frame, host arrays in, host draw list out (median of %d frames)" % frames}
    if not cpu_reference:
        return res
    try:
        # CPU-baseline leg (the one place bench.py may execute oracle/): the reference's own BillboardSystem.cpp
        from oracle import billboards as B
        if B.have_reference():
            o = B.BillboardOracle()
            ct = []
            for _ in range(3):
                t0 = time.perf_counter()
                o.lib.orc_bb_add(o.h, ptr, n)
                # the reference cuts the frame at 4096 billboards (:10, :121-123): its sort still runs over all n
                q = o.lib.orc_bb_render(o.h, C.byref(ra), None, 0, None, 0, None, None)
                ct.append(time.perf_counter() - t0)
            o.close()
            res["cpu_reference"] = {"ms_per_frame": sorted(ct)[1] * 1e3, "quads": int(q), "cores": 1, "kind": "reference",
                                    "note": "oracle/_ref: the reference's BillboardSystem.cpp on the same frame (sorts all n, expands its budget of 4096)"}
    except Exception as ex:  # noqa: BLE001
        res["cpu_reference"] = {"error": str(ex)[:200]}
    return res


# ---------------------------------------------------------------------------
# CPU arms (oracle/_ref = the reference's own code; never the product path)
# ---------------------------------------------------------------------------

def _cpu_worker(kind, name, idx, steps, warmup, scale, q):
    from oracle import orc
    from spear_b200 import abi
    wl = workload_components(name)
    if "effects" in wl:
        o = orc.OracleSystem(kind, (1, 2, 3, 4))
        effs = wl["effects"]
        if scale < 1.0 and len(effs) == 1:
            effs[0][0].burstAmount = int(effs[0][0].burstAmount * scale)
        ids = [o.add_effect(c, t) for c, t in effs]
    else:
        o = orc.OracleSystem(kind, (1, 2 + idx, 3, 4))
        ids = [o.add_effect(wl["comps"][idx % len(wl["comps"])], abi.make_transform((idx + 0.25 * k, 0, 0)))
               for k in range(wl.get("effects_per_system", 1))]
    dt = max(o.effect_info(e).timeStep for e in ids)
    fa = abi.make_frame_args(dt, camera=CAMERA)
    o.bench_steps(ids, fa, 2 + warmup)          # burst + warm-up, untimed
    s0 = sum(o.effect_info(e).simSteps for e in ids)
    secs, _ = o.bench_steps(ids, fa, steps)
    s1 = sum(o.effect_info(e).simSteps for e in ids)
    alive = sum(o.effect_info(e).aliveCount for e in ids)
    q.put((secs, (s1 - s0) * alive / len(ids)))


def cpu_run(name, steps, warmup, procs, scale=1.0):
    import multiprocessing as mp
    from oracle import orc
    kind = "reference" if orc.have("reference") else "port"
    if kind == "port":
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    # instance 0 runs in THIS process (the reference library is loaded here, where the driver's loaded-library record
    # looks), the other instances in forked children of it
    orc._load(kind)
    ctxm = mp.get_context("fork")
    q = ctxm.Queue()
    ps = [ctxm.Process(target=_cpu_worker, args=(kind, name, i, steps, warmup, scale, q)) for i in range(1, procs)]
    t0 = time.perf_counter()
    for p in ps:
        p.start()
    import queue as _queue
    local = _queue.Queue()
    _cpu_worker(kind, name, 0, steps, warmup, scale, local)
    res = [local.get()] + [q.get() for _ in ps]
    for p in ps:
        p.join()
    wall = time.perf_counter() - t0
    secs = max(r[0] for r in res)
    total = sum(r[1] for r in res)
    return kind, total / secs, secs, wall


def cpu_baseline(name, seconds=15.0):
    """Bounded sample of the same workload on the host cores (rank 0, N=1)."""
    if name == "c3":
        kind, v, secs, wall = cpu_run("c3", steps=20, warmup=1, procs=1)
        sample = "the full C3 effect (10M particles) on 1 core: burst + 1 warm-up step untimed, 20 timed steps; the reference has no per-particle sort, so this is update+compact+4x expand"
        cores = 1
    elif name == "c2":
        kind, v, secs, wall = cpu_run("c2", steps=100, warmup=3, procs=1)
        sample = "the full C2 system (64 x 16k) on 1 core (one system shares one RNG, so it cannot be threaded), 100 timed steps"
        cores = 1
    else:
        cores = os.cpu_count() or 1
        kind, v, secs, wall = cpu_run(name, steps=60, warmup=6, procs=cores)
        sample = "%d of the per-GPU system instances of %s, one reference system per host core, 60 timed steps" % (cores, name)
    return {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "seconds": secs}


def run_reference(args, rank, world):
    if rank != 0:
        return
    name = args.workload or ("c3" if world == 1 else "c4")
    wl = workload_components(name)
    cores_avail = os.cpu_count() or 1
    t0 = time.perf_counter()
    if name in ("c3", "c2"):
        # a single system is single-threaded in the reference (one shared RNG, ParticleSystem.cpp:251)
        kind, v, secs, wall = cpu_run(name, steps=args.steps, warmup=args.warmup, procs=1)
        cores = 1
        sample = "%s on 1 host core (the path is single-threaded per system): %d timed steps after burst + %d warm-up" % (wl["desc"], args.steps, args.warmup)
    else:
        cores = cores_avail
        kind, v, secs, wall = cpu_run(name, steps=args.steps, warmup=args.warmup, procs=cores)
        sample = "%d instances of %s, one per host core, %d timed steps" % (cores, wl["desc"], args.steps)
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "name": name, "note": "reference CPU ParticleSystem (no per-particle sort exists in the reference: update+compact+4x expand)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--repeats", type=int, default=5, help="timed regions of --steps steps each; value = the median region")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "c2", "c3", "c4", "c5"])
    ap.add_argument("--no-sort", action="store_true", help="strict drop-in stream (slot order), no per-particle depth sort")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stream-mode", action="store_true", help="skip the extra stream-mode (no sort) measurement at N=1")
    ap.add_argument("--no-billboards", action="store_true", help="skip the extra cl::BillboardSystem (row N4) measurement at N=1")
    ap.add_argument("--no-vertex-stage", action="store_true", help="skip the extra vertex-stage (sprShadeVertices) measurement at N=1")
    ap.add_argument("--assembly-shares", default="1,2,3,8", help="N > 1: CTAs per SM of the peer-memory assembly kernel to try in the overlapped run (best one is reported)")
    ap.add_argument("--no-c4-slice", action="store_true", help="skip the extra C4-slice measurement (the N > 1 per-GPU workload) at N=1")
    ap.add_argument("--no-kernel-timing", action="store_true", help="no per-launch CUDA events inside the timed region (no roofline object)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    # a run that stops making progress dumps every thread's stack to stderr and exits instead of sitting on the GPU box
    import faulthandler
    faulthandler.enable()
    faulthandler.dump_traceback_later(int(os.environ.get("SPR_BENCH_WATCHDOG_S", "900")), exit=True)
    main()
