#!/usr/bin/env python
"""Benchmark of the hot path: one STEP = one frame of the workload through the ray-query path
(primary rays + one coherent shadow-ray queue per light), scene resident in HBM.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload dragon4k] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU): the scene is replicated, image tiles are interleaved over ranks
(dodrt_frame first_tile/tile_stride), every rank traces its tiles with ONE launch (dodrt_trace_frame_device) whose
kernel also stores each result straight into rank 0's row-major frame buffer over NVLink (dodrt_frame_buffer_*, CUDA
IPC): no gather, no assembly pass.  torch.distributed (NCCL) only carries the 96-byte frame-buffer descriptor, the
barriers and the max-over-ranks of the timings.  The frame is fixed, so this is STRONG scaling.

Prints ONE JSON line (rank 0).  `value` = Mrays/s (primary + shadow) with inputs resident in HBM, device-timed with
CUDA events on the launching stream, max over ranks; `e2e` = the same metric through the host-buffer C-ABI call
(dodrt_trace_frame: H2D of the ray tables inside, every hit record and visibility byte delivered into pinned HOST
memory inside the timed region).  `roofline` leads with the resource that binds this kernel -- instruction issue --
and carries the HBM view (compulsory and measured bytes, and the per-ray algorithmic figure of SURVEY 8(d) as context);
`cpu_baseline` is the reference's own CPU code (oracle/_ref) or, when that library did not travel, the oracle port,
timed on this box's host cores.  `--impl reference` times that CPU implementation as the arm itself, without mapping
any of the product's libraries.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "Mrays/s (primary+shadow)"
UNIT = "Mrays/s"
FRAME_KERNEL = "frame share (primary + shadow passes)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dragon4k")
    ap.add_argument("--tile", default="32x32")
    ap.add_argument("--separate", action="store_true", help="N > 1: issue the passes one by one (per-pass event times) instead of dodrt_trace_frame_device")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N > 1: ranks store into rank 0's frame over NVLink (peer) or NCCL gather + assembly kernel (A/B, fallback)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-render", action="store_true", help="skip the reference's as-is frame (9 lights, 10 bounces)")
    return ap.parse_args()


def config_of(w, args):
    """The workload as BOTH arms print it (identical dicts: same config, same metric)."""
    from dod_raytracer_b200 import workloads
    return {"workload": w.name, "description": w.description, "mesh": workloads.mesh_label(w), "width": w.width,
            "height": w.height, "primary_rays": w.pixels, "lights": len(w.lights) if w.shadow else 0, "classes": w.classes,
            "l2": "GPU arm: flushed between steps (512 MiB memset outside the per-step event bracket); reference arm: CPU caches as they come"}


# ---- clocks ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md 'clocks line')."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU side: the reference's own implementation of the path ------------------------------------------------
class quiet_stdout:
    """The reference printf()s progress lines (kdtree.cpp:255-257, triangle.cpp:356,366); keep them out of
    this program's stdout, which must carry exactly one JSON line.  Redirects fd 1 at the OS level and
    flushes C stdio before restoring it."""

    def __enter__(self):
        import ctypes
        sys.stdout.flush()
        self._libc = ctypes.CDLL(None)
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)
        return self

    def __exit__(self, *exc):
        self._libc.fflush(None)
        os.dup2(self._saved, 1)
        os.close(self._saved)
        os.close(self._null)


class CpuPath:
    """oracle/_ref (the reference's translation units, kind 'reference') when it travelled, else the oracle
    restatement (kind 'port').  Only used for cpu_baseline, the parity check and --impl reference.  `host_arrays_fn`
    (the product's host builder) is only called by the port when a kd-tree is needed and _ref is absent."""

    def __init__(self, w, mesh_files, host_arrays_fn=None):
        import oracle_api
        self.api, self.w = oracle_api, w
        self.cores = os.cpu_count() or 1
        self.light = np.array(w.lights[0], np.float32)
        if oracle_api.have_ref() and w.mesh != "none" and not (w.classes & 16):
            self.kind = "reference"
            with quiet_stdout():
                ref = oracle_api.RefLib()
                ref.set_config(w.width, w.height)
                if w.reference_scene:
                    ref.add_reference_spheres(1, 16)
                    ref.add_reference_planes()
                    ref.add_reference_cylinder()
                for path in mesh_files:
                    ref.add_mesh(path)
                ref.build_tree()
            self.ref = ref
        else:
            self.kind = "port"
            oracle_api.ensure_oracle_built()
            self.orc = oracle_api.Oracle()
            if w.mesh == "none" and w.analytic:  # config 4: no kd-tree, no product library needed
                from scenes import analytic_scene_arrays
                spheres, boxes = analytic_scene_arrays(4, w.analytic)
                self.scene = oracle_api.Scene(spheres=spheres, boxes=boxes)
            else:
                self.scene = oracle_api.Scene.from_host_arrays(host_arrays_fn())

    def frame(self):
        """one full frame on all host cores: returns (seconds, rays, t [n] float32, visible [n] uint8)"""
        w = self.w
        t0 = time.perf_counter()
        if self.kind == "reference":
            with quiet_stdout():
                t, vis = self.ref.trace_frame(w.width, w.height, w.classes & 15, self.light, self.cores)
            if not w.shadow:
                vis[:] = 0
            hit = np.isfinite(t)
        else:
            hits = self.orc.trace_primary(self.scene, w.width, w.height, w.classes, nthreads=self.cores)
            hit = hits["prim"] != 0xFFFFFFFF
            t = np.where(hit, hits["t"], np.float32(np.inf)).astype(np.float32)
            vis = (self.orc.trace_shadow(self.scene, w.width, w.height, w.classes, hits, self.light, nthreads=self.cores)
                   if w.shadow else np.zeros(w.pixels, np.uint8))
        dt = time.perf_counter() - t0
        rays = w.pixels + (int(hit.sum()) if w.shadow else 0)
        return dt, rays, t, vis

    def describe(self, seconds, rays):
        return {"value": rays / seconds / 1e6, "unit": UNIT, "cores": self.cores, "kind": self.kind,
                "sample": f"the whole workload: 1 frame = {rays} rays in {seconds * 1e3:.0f} ms, best of 2, row bands "
                          f"over {self.cores} threads (main.cpp:371-393), g++ -O2 -ffp-contract=off -mavx2",
                "frame_ms": seconds * 1e3}


def parity_of(cpu_t, cpu_vis, gpu_hits, gpu_vis, shadow, kind, rays):
    from dod_raytracer_b200 import capi
    t_gpu = np.where(gpu_hits["prim"] != capi.MISS, gpu_hits["t"], np.float32(np.inf)).astype(np.float32)
    return {"against": kind, "t_bit_mismatches": int((t_gpu.view(np.uint32) != cpu_t.view(np.uint32)).sum()),
            "visibility_mismatches": int((gpu_vis != cpu_vis).sum()) if shadow else 0, "rays": int(rays)}


def run_reference_arm(args, w):
    """The reference's CPU implementation as the arm: no product library is mapped into this process (meshes come from
    the pure-Python generator in tests/scenes.py, byte-identical to the host library's)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from scenes import workload_mesh_files_py
    tmpdir = tempfile.mkdtemp(prefix="dodrt_bench_ref_")
    mesh_files = workload_mesh_files_py(w, tmpdir)

    def host_arrays():  # port fallback over a kd-tree only (no _ref on this box): needs the product's host builder
        from dod_raytracer_b200 import workloads
        return workloads.build_host_scene(w, mesh_files).arrays()

    cpu = CpuPath(w, mesh_files, host_arrays)
    for _ in range(args.warmup):
        cpu.frame()
    total, rays = 0.0, 0
    for _ in range(args.steps):
        dt, r, _, _ = cpu.frame()
        total += dt
        rays += r
    value = rays / total / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(w, args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind,
                             "sample": f"each step = the whole frame ({rays // max(args.steps, 1)} rays) on {cpu.cores} host threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- roofline ----------------------------------------------------------------------------------------------------
def roofline_of(w, world, shadow_rays, kernel_ms, per_pass_ms, sizes, nl, clocks, separate):
    """What binds the frame kernel, and how far from it we are.

    * issue (primary entry): warp-instructions per launch (ncu smsp__inst_executed.sum of the same kernel on the same
      workload, profiles/traffic.json) / the LIVE launch time (CUDA events here) against 148 SMs x 4 schedulers x the SM
      clock sampled during the run; `active_lanes` says how many of the 32 lanes an issued instruction used.
    * hbm: the same launch time against (a) the compulsory bytes -- every node and lane of the scene once + results
      out + hit records read back by the shadow queue -- and (b) the measured dram__bytes of the ncu capture.
    * per_ray_cache_served: SURVEY 8(d)'s algorithmic figure (8 B/node + 288 B/lane + io per ray of the REFERENCE
      traversal).  32 coherent rays share one fetch through L1/L2, so this is a demand on the caches, not on HBM; it is
      context, not a roofline (it exceeds the HBM peak)."""
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    b_prim, b_shad = w.algorithmic_bytes(shadow_rays)
    algorithmic = (b_prim + b_shad * nl) / world
    scene_bytes = 8.0 * sizes.num_nodes + 288.0 * sizes.num_lanes + 128.0 * ((sizes.num_spheres + 7) // 8) + \
        192.0 * ((sizes.num_boxes + 7) // 8)
    io_bytes = (16.0 * w.pixels + (16.0 + 1.0) * shadow_rays * nl) / world
    compulsory = scene_bytes + io_bytes
    sec = kernel_ms * 1e-3
    hbm = {"compulsory_bytes_per_launch": compulsory, "compulsory_GBs": compulsory / sec / 1e9 if sec else None,
           "compulsory_frac": compulsory / sec / 1e9 / peak if sec else None, "peak": peak, "peak_source": peak_src,
           "measured_dram_bytes_per_launch": None,
           "per_ray_cache_served": {"bytes_per_launch": algorithmic, "GBs": algorithmic / sec / 1e9 if sec else None,
                                    "note": "8 B/node + 288 B/lane + io per ray of the reference traversal (oracle-counted); "
                                            "served by L1/L2 to 32 coherent rays at a time, not by HBM"}}
    name = FRAME_KERNEL if not separate else ("trace_kernel<shadow>" if (nl and per_pass_ms[1] >= per_pass_ms[0]) else "trace_kernel<primary>")
    dom_ms = kernel_ms if not separate else (per_pass_ms[1] if name.endswith("<shadow>") else per_pass_ms[0])
    roof = {"bound": "hbm", "kernel": name, "achieved": hbm["compulsory_GBs"], "peak": peak, "unit": "GB/s",
            "frac": hbm["compulsory_frac"], "traffic": None, "kernel_ms": dom_ms, "hbm": hbm,
            "note": "no ncu counters for this workload/kernel in profiles/traffic.json: HBM view against compulsory bytes only"}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    hw = None
    if os.path.exists(prof):
        try:  # N > 1: the counters of rank 0's share of an N-way split, when that split was captured
            hw = json.load(open(prof)).get(w.name if world == 1 else f"{w.name}@{world}", {}).get(name)
        except Exception:
            hw = None
    if hw:
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        issue_peak = 148 * 4 * sm_mhz * 1e6 / 1e9  # G warp-instructions / s
        achieved = hw["warp_inst"] / (dom_ms * 1e-3) / 1e9
        hbm["measured_dram_bytes_per_launch"] = hw["dram_bytes"]
        hbm["measured_GBs"] = hw["dram_bytes"] / (dom_ms * 1e-3) / 1e9
        hbm["measured_frac"] = hbm["measured_GBs"] / peak
        roof.update({"bound": "issue", "achieved": achieved, "peak": issue_peak, "unit": "Gwarp-inst/s", "frac": achieved / issue_peak,
                     "traffic": hw["dram_bytes"], "active_lanes": hw["lanes_per_inst"],
                     "lane_frac": achieved / issue_peak * hw["lanes_per_inst"] / 32.0,
                     "warp_inst_per_launch": hw["warp_inst"], "sm_mhz": sm_mhz,
                     "note": "bound = instruction issue: warp-instructions per launch from the ncu capture of this kernel on this "
                             "workload (profiles/traffic.json, " + str(hw.get("source", "profiles/")) + ") over the live CUDA-event "
                             "launch time, against 148 SMs x 4 issue slots x the sampled SM clock; lane_frac = frac x active_lanes/32. "
                             "The HBM view is under `hbm`: DRAM traffic is ~1 % of peak, the kernel is not memory bound"})
    return roof


# ---- the GPU arm ---------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    from dod_raytracer_b200 import workloads
    w = workloads.WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, w)
        return

    from dod_raytracer_b200 import capi, distributed, host
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    tmpdir = tempfile.mkdtemp(prefix=f"dodrt_bench_{rank}_")
    mesh_files = workloads.write_mesh_files(w, tmpdir)
    t0 = time.perf_counter()
    # lanes stay in creation order, Triangle::reorderLanesByIndices runs on the GPU at upload (f-4)
    hs = workloads.build_host_scene(w, mesh_files, keep_creation_order=True)
    build_s = time.perf_counter() - t0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    if world != args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch N>1 with torchrun")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    t_up = time.perf_counter()
    scene = hs.upload(local_rank)  # lanes in creation order + m_primNums: the GPU does the lane re-order (f-4)
    torch.cuda.synchronize()
    upload_s = time.perf_counter() - t_up
    sizes = hs.sizes()
    tile = tuple(int(x) for x in args.tile.split("x"))
    frame = distributed.rank_frame(w.width, w.height, w.classes, rank, world, tile)
    slots = capi.frame_local_pixels(frame) if world > 1 else w.pixels
    slots_rank0 = distributed.slots_per_rank(w.width, w.height, world, tile)
    xs, ys = host.ray_tables(w.width, w.height)
    d_xs, d_ys = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
    lights = np.array(w.lights, np.float32)
    nl = len(lights) if w.shadow else 0
    # per-rank result buffers (padded to rank 0's slot count so an NCCL gather, if used, is regular)
    d_hits = torch.empty((slots_rank0 if world > 1 else slots, 16), dtype=torch.uint8, device=dev)
    d_vis = torch.zeros((max(nl, 1), slots_rank0 if world > 1 else slots), dtype=torch.uint8, device=dev)
    if world > 1:
        d_hits.fill_(0xFF)

    # ---- N > 1: rank 0's frame buffer, opened by every other rank over CUDA IPC ------------------------------------
    mirror, owner_fb, gather = None, None, args.gather
    if world > 1 and gather == "peer":
        ok = 1
        desc_t = torch.zeros(96, dtype=torch.uint8, device=dev)
        try:
            if rank == 0:
                owner_fb = capi.FrameBuffer.create(scene, w.width, w.height, nl)
                desc_t.copy_(torch.frombuffer(bytearray(owner_fb.export().to_bytes()), dtype=torch.uint8))
        except Exception as exc:  # noqa: BLE001
            print(f"bench.py: frame buffer export failed on rank 0: {exc!r}", file=sys.stderr)
            ok = 0
        dist.broadcast(desc_t, src=0)
        try:
            if rank == 0:
                mirror = owner_fb
            elif ok:
                mirror = capi.FrameBuffer.open(scene, capi.FrameBufferDesc.from_bytes(bytes(desc_t.cpu().numpy())))
        except Exception as exc:  # noqa: BLE001
            print(f"bench.py: rank {rank} cannot open rank 0's frame buffer: {exc!r}", file=sys.stderr)
            ok = 0
        okt = torch.tensor([ok if mirror is not None else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if int(okt.item()) == 0:
            gather, mirror = "nccl", None  # every rank falls back together
    if world > 1 and gather == "nccl" and rank == 0:
        g_hits = torch.empty((world, slots_rank0, 16), dtype=torch.uint8, device=dev)
        g_vis = torch.empty((world, slots_rank0), dtype=torch.uint8, device=dev)
        f_hits = torch.empty((w.pixels, 16), dtype=torch.uint8, device=dev)
        f_vis = torch.empty(w.pixels, dtype=torch.uint8, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=dev) if (world > 1 and gather == "nccl") else None  # carries the hit-record gather
    hits_ready, hits_gathered = torch.cuda.Event(), torch.cuda.Event()
    # N = 1: the two passes of dodrt_trace_frame_device are issued one by one here so that CUDA events can time each
    # kernel on its own (same kernels, same order); N > 1: the one call a rank makes, results mirrored into rank 0's frame
    two_pass = world == 1 or args.separate or gather == "nccl"

    def step(ev=None):
        sp = stream.cuda_stream
        if ev:
            ev[0].record(stream)
        if not two_pass:
            # THE product path: one launch for the rank's share, results mirrored into rank 0's frame when N > 1
            scene.trace_frame_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), lights[:nl], d_hits.data_ptr(), d_vis.data_ptr(),
                                     mirror, sp)
            if ev:
                ev[1].record(stream)
                ev[2].record(stream)
        else:
            scene.trace_primary_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), sp)
            if ev:
                ev[1].record(stream)
            if side is not None:
                hits_ready.record(stream)
                with torch.cuda.stream(side):
                    side.wait_event(hits_ready)
                    distributed.gather_to_rank0(d_hits, world, rank, g_hits if rank == 0 else None)
                    hits_gathered.record(side)
            for l in range(nl):
                scene.trace_shadow_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), lights[l],
                                          d_vis[l].data_ptr(), sp)
            if ev:
                ev[2].record(stream)
            if side is not None:
                distributed.gather_to_rank0(d_vis[0], world, rank, g_vis if rank == 0 else None)
                stream.wait_event(hits_gathered)
                if rank == 0:
                    scene.frame_assemble_device(frame, g_hits.data_ptr(), g_vis.data_ptr(), slots_rank0, f_hits.data_ptr(),
                                                f_vis.data_ptr(), sp)
        if ev:
            ev[3].record(stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step()
    sync_all()

    def assembled_frame():
        """rank 0: the row-major frame as the run left it (hits [pixels], vis [pixels])"""
        if world == 1:
            return d_hits.cpu().numpy().reshape(-1).view(capi.HIT_DT), d_vis[0].cpu().numpy()
        if gather == "nccl":
            return f_hits.cpu().numpy().reshape(-1).view(capi.HIT_DT), f_vis.cpu().numpy()
        hp, vp = owner_fb.pointers()

        class Raw:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        fh = torch.as_tensor(Raw(hp, w.pixels * 16), device=dev).cpu().numpy().view(capi.HIT_DT)
        fv = torch.as_tensor(Raw(vp, w.pixels), device=dev).cpu().numpy() if nl else np.zeros(w.pixels, np.uint8)
        return fh, fv

    local_hits = d_hits.cpu().numpy().reshape(-1).view(capi.HIT_DT)
    n_hit = torch.tensor([int((local_hits["prim"] != capi.MISS).sum())], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(n_hit)
    shadow_rays = int(n_hit.item()) if w.shadow else 0
    rays_per_step = w.pixels + shadow_rays * nl

    # ---- timed region: exactly K steps, barrier + synchronize on both sides, device-timed per step ----------
    events = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    launches0 = scene.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sync_all()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()  # L2 flush between steps (outside the per-step event bracket)
        step(events[k])
    sync_all()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = scene.launch_count() - launches0
    step_ms = [e[0].elapsed_time(e[3]) for e in events]
    prim_ms = [e[0].elapsed_time(e[1]) for e in events]
    shad_ms = [e[1].elapsed_time(e[2]) for e in events]
    total_ms = torch.tensor([sum(step_ms), sum(prim_ms), sum(shad_ms)], device=dev, dtype=torch.float64)
    rank_ms = [sum(step_ms) / args.steps]
    if world > 1:
        every = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(every, total_ms[:1].clone())
        rank_ms = [float(t.item()) / args.steps for t in every]
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    total_ms, prim_total, shad_total = [float(x) for x in total_ms.tolist()]
    ms_per_step = total_ms / args.steps
    value = rays_per_step / (ms_per_step * 1e-3) / 1e6
    res_hits, res_vis = assembled_frame() if rank == 0 else (None, None)

    # ---- e2e: host buffers through the C ABI, every rank its own tiles; the call returns when every record is in the
    # caller's (pinned) host buffers.
    e2e = None
    if not args.no_e2e:
        h_hits = torch.empty((slots, 16), dtype=torch.uint8, pin_memory=True).numpy().reshape(-1).view(capi.HIT_DT)
        h_vis = torch.empty((max(nl, 1), slots), dtype=torch.uint8, pin_memory=True).numpy()
        h_xs = torch.from_numpy(xs).pin_memory().numpy()
        h_ys = torch.from_numpy(ys).pin_memory().numpy()
        for _ in range(3):
            scene.trace_frame(frame, h_xs, h_ys, lights[:nl], h_hits, h_vis)
        sync_all()
        t_e2e = 0.0
        for _ in range(args.steps):
            flush.zero_()
            sync_all()
            t1 = time.perf_counter()
            scene.trace_frame(frame, h_xs, h_ys, lights[:nl], h_hits, h_vis)  # synchronous: returns when the host has it all
            t_e2e += time.perf_counter() - t1
        te = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item()) / args.steps
        same = bool(h_hits[:slots].tobytes() == local_hits[:slots].tobytes())
        e2e = {"value": rays_per_step / e2e_s / 1e6, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": int(xs.nbytes + ys.nbytes) * world,
               "d2h_bytes_per_step": int(w.pixels * (16 + nl)) if world == 1 else int(slots_rank0 * (16 + nl)) * world,
               "api": "dodrt_trace_frame (pinned host buffers; persistent staging, hit-record D2H overlapped with the shadow pass)",
               "steps": args.steps, "host_results_equal_device_results": same}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    roofline = roofline_of(w, world, shadow_rays, ms_per_step, (prim_total / args.steps, shad_total / args.steps), sizes, nl,
                           clocks, two_pass)

    # ---- CPU baseline + parity of the ASSEMBLED frame (outside every timed region), at every N --------------------------
    cpu_baseline, parity, cpu = None, None, None
    if not args.no_cpu_baseline:
        cpu = CpuPath(w, mesh_files, hs.arrays)
        best = None
        for _ in range(2 if world == 1 else 1):
            dt, r, t_cpu, vis_cpu = cpu.frame()
            best = dt if best is None else min(best, dt)
        if world == 1:
            cpu_baseline = cpu.describe(best, r)
        parity = parity_of(t_cpu, vis_cpu, res_hits, res_vis if nl else np.zeros(w.pixels, np.uint8), nl > 0, cpu.kind, r)
        parity["frame"] = "one GPU" if world == 1 else f"assembled from {world} ranks ({gather})"

    # ---- the reference's as-is frame (rayTrace, main.cpp:273-347: 9 lights, 10 bounces) at config.ini's 1920x1080 ------
    reference_frame = None
    if world == 1 and not args.no_render and not args.no_cpu_baseline and w.reference_scene and w.mesh != "none":
        try:
            rw, rh = 1920, 1080
            rscene = hs.upload(local_rank, shading=True)
            rxs, rys = host.ray_tables(rw, rh)
            rframe = capi.Frame.make(rw, rh, classes=workloads.CLS_REFERENCE)
            rgb = torch.empty((rh, rw, 3), dtype=torch.uint8, pin_memory=True).numpy()
            rscene.render(rframe, rxs, rys, workloads.REFERENCE_LIGHTS, workloads.REFERENCE_DEPTH, rgb)
            best = None
            for _ in range(3):
                t1 = time.perf_counter()
                rscene.render(rframe, rxs, rys, workloads.REFERENCE_LIGHTS, workloads.REFERENCE_DEPTH, rgb)
                dt = time.perf_counter() - t1
                best = dt if best is None else min(best, dt)
            rscene.close()
            queries = rw * rh * workloads.REFERENCE_DEPTH * (1 + len(workloads.REFERENCE_LIGHTS))
            reference_frame = {"what": "rayTrace as the reference ships it: 1920x1080, 9 lights, 10 mirror bounces, all shape "
                                       "classes, 8-bit RGB to the host (dodrt_render)", "ray_queries": queries,
                               "gpu_frame_ms": best * 1e3, "gpu_mrays_s": queries / best / 1e6}
            if cpu.kind == "reference":
                with quiet_stdout():
                    t1 = time.perf_counter()
                    cpu.ref.render(rw, rh, nthreads=cpu.cores)
                    cpu_s = time.perf_counter() - t1
                reference_frame.update(cpu_frame_ms=cpu_s * 1e3, cpu_cores=cpu.cores, speedup=cpu_s / best)
        except Exception as exc:  # never lose the headline line over the extra
            reference_frame = {"error": repr(exc)}

    if world == 1:
        how = "primary pass + one coherent shadow pass per light (persistent kernels, warp-voted traversal)"
    elif gather == "peer":
        how = (f"image tiles round-robin over {world} GPUs, scene replicated; dodrt_trace_frame_device per rank (primary + shadow "
               "pass, donating kernels), every result stored straight into rank 0's row-major frame buffer over NVLink by the "
               "kernels themselves (CUDA IPC peer mapping); no gather, no assembly pass")
    else:
        how = (f"image tiles round-robin over {world} GPUs, scene replicated; NCCL gather to rank 0 overlapped with the shadow pass + "
               "dodrt_frame_assemble_device")
    standin = w.mesh.startswith("dragon") and not os.environ.get("DODRT_DRAGON_OBJ")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_of(w, args),
            "details": {"shadow_rays": shadow_rays * nl, "triangles": sizes.num_triangles, "kd_nodes": sizes.num_nodes,
                        "tri_lanes": sizes.num_lanes, "tile": args.tile, "parallelism": how,
                        "host_build_s": round(build_s, 2), "upload_s": round(upload_s, 2), "wall_s_timed_region": round(wall, 3),
                        "rank_ms_per_step": [round(x, 4) for x in rank_ms],
                        "mesh_note": "stand-in geometry: assets/dragon.obj is a missing blob in the reference" if standin else None},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "parity": parity, "frame_ms": ms_per_step, "reference_frame": reference_frame}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
