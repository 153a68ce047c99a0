#!/usr/bin/env python
"""Benchmark of the hot path: one STEP = one frame of the workload through the ray-query path
(primary rays, then one coherent shadow-ray pass per light), scene resident in HBM.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload dragon4k] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU): the scene is replicated, image tiles are interleaved
over ranks (dodrt_frame first_tile/tile_stride), every rank traces its tiles into a compact buffer,
the buffers are gathered on rank 0 over NVLink (NCCL) and re-assembled into the row-major frame.
The frame is fixed, so this is STRONG scaling.

Prints ONE JSON line (rank 0).  `value` = Mrays/s (primary + shadow) with inputs resident in HBM,
device-timed with CUDA events on the launching stream, max over ranks; `e2e` = the same metric through the
host-buffer C-ABI call (dodrt_trace_frame: H2D of the ray tables, D2H of hit records + visibility inside
the timed region).  `roofline` is for the dominant kernel against the measured HBM peak using the
reference traversal's algorithmic bytes (dod_raytracer_b200/algorithmic_bytes.json); `cpu_baseline` is
the reference's own CPU code (oracle/_ref) or, when that library did not travel, the oracle port, timed
on this box's host cores.  `--impl reference` times that CPU implementation as the arm itself.
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

import numpy as np  # noqa: E402

METRIC = "Mrays/s (primary+shadow)"
UNIT = "Mrays/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dragon4k")
    ap.add_argument("--tile", default="32x32")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-render", action="store_true", help="skip the reference's as-is frame (9 lights, 10 bounces)")
    return ap.parse_args()


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
    restatement (kind 'port').  Only used for cpu_baseline and --impl reference."""

    def __init__(self, w, mesh_files, host_arrays_fn):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
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
            t = hits["t"]
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


def run_reference_arm(args, w, mesh_files, host_arrays_fn):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu = CpuPath(w, mesh_files, host_arrays_fn)
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
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w.name, "description": w.description, "width": w.width, "height": w.height},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind,
                             "sample": f"each step = the whole frame ({rays // args.steps} rays) on {cpu.cores} host threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- the GPU arm ---------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    from dod_raytracer_b200 import capi, distributed, host, workloads
    w = workloads.WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    tmpdir = tempfile.mkdtemp(prefix=f"dodrt_bench_{rank}_")
    mesh_files = workloads.write_mesh_files(w, tmpdir)
    t0 = time.perf_counter()
    # GPU arm: lanes stay in creation order, Triangle::reorderLanesByIndices runs on the GPU at upload (f-4)
    hs = workloads.build_host_scene(w, mesh_files, keep_creation_order=(args.impl != "reference"))
    build_s = time.perf_counter() - t0

    if args.impl == "reference":
        run_reference_arm(args, w, mesh_files, hs.arrays)
        return

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
    # per-rank result buffers (padded to rank 0's slot count so the gather is regular)
    d_hits = torch.empty((slots_rank0 if world > 1 else slots, 16), dtype=torch.uint8, device=dev)
    d_vis = torch.zeros((max(nl, 1), slots_rank0 if world > 1 else slots), dtype=torch.uint8, device=dev)
    if world > 1:
        d_hits.fill_(0xFF)
    if world > 1 and rank == 0:
        g_hits = torch.empty((world, slots_rank0, 16), dtype=torch.uint8, device=dev)
        g_vis = torch.empty((world, slots_rank0), dtype=torch.uint8, device=dev)
        f_hits = torch.empty((w.pixels, 16), dtype=torch.uint8, device=dev)
        f_vis = torch.empty(w.pixels, dtype=torch.uint8, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    side = torch.cuda.Stream(device=dev) if world > 1 else None  # carries the hit-record gather
    hits_ready, hits_gathered = torch.cuda.Event(), torch.cuda.Event()

    def step(ev=None):
        sp = stream.cuda_stream
        if ev:
            ev[0].record(stream)
        scene.trace_primary_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), sp)
        if ev:
            ev[1].record(stream)
        if world > 1:
            # 94 % of the gathered bytes are the primary hit records: ship them over NVLink on a side stream
            # WHILE the shadow pass (which only reads them) runs on the main stream
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
        if world > 1:
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
    res_hits = (f_hits if (world > 1 and rank == 0) else d_hits).cpu().numpy().reshape(-1).view(capi.HIT_DT)
    if world > 1:
        n_hit_local = torch.tensor([int((d_hits.cpu().numpy().reshape(-1).view(capi.HIT_DT)["prim"] != capi.MISS).sum())],
                                   device=dev, dtype=torch.int64)
        dist.all_reduce(n_hit_local)
        shadow_rays = int(n_hit_local.item()) if w.shadow else 0
    else:
        shadow_rays = int((res_hits["prim"] != capi.MISS).sum()) if w.shadow else 0
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
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    total_ms, prim_total, shad_total = [float(x) for x in total_ms.tolist()]
    ms_per_step = total_ms / args.steps
    value = rays_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host buffers through the C ABI (H2D + D2H inside the timed region), every rank its own tiles -------
    e2e = None
    if not args.no_e2e:
        h_hits = torch.empty((slots, 16), dtype=torch.uint8, pin_memory=True).numpy().reshape(-1).view(capi.HIT_DT)
        h_vis = torch.empty((max(nl, 1), slots), dtype=torch.uint8, pin_memory=True).numpy()
        h_xs = torch.from_numpy(xs).pin_memory().numpy()
        h_ys = torch.from_numpy(ys).pin_memory().numpy()
        e2e_steps = min(args.steps, 5)
        for _ in range(2):
            scene.trace_frame(frame, h_xs, h_ys, lights[:nl], h_hits, h_vis)
        sync_all()
        t_e2e = 0.0
        for _ in range(e2e_steps):
            flush.zero_()
            sync_all()
            t1 = time.perf_counter()
            scene.trace_frame(frame, h_xs, h_ys, lights[:nl], h_hits, h_vis)  # synchronous: returns after D2H
            t_e2e += time.perf_counter() - t1
        te = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item()) / e2e_steps
        e2e = {"value": rays_per_step / e2e_s / 1e6, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": int(xs.nbytes + ys.nbytes) * world,
               "d2h_bytes_per_step": int(w.pixels * (16 + nl)) if world == 1 else int(slots_rank0 * (16 + nl)) * world,
               "api": "dodrt_trace_frame (host buffers, pinned)", "steps": e2e_steps}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------------
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    b_prim, b_shad = w.algorithmic_bytes(shadow_rays)
    b_prim, b_shad = b_prim / world, b_shad / world  # per launch = this rank's tiles (even split assumed)
    k_prim, k_shad = prim_total / args.steps, shad_total / args.steps
    dominant = "trace_kernel<shadow>" if (nl and k_shad >= k_prim) else "trace_kernel<primary>"
    dom_bytes, dom_ms = (b_shad, k_shad) if dominant.endswith("<shadow>") else (b_prim, k_prim)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms,
                "kernels": {"primary": {"ms": k_prim, "GB/s": b_prim / (k_prim * 1e-3) / 1e9 if k_prim else None},
                            "shadow": {"ms": k_shad, "GB/s": b_shad / (k_shad * 1e-3) / 1e9 if k_shad else None}},
                "note": "algorithmic bytes = reference traversal's 8 B/node + 288 B/lane + io (oracle-counted); "
                        "traffic (ncu dram bytes) is in profiles/"}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof) and world == 1:
        try:
            hw = json.load(open(prof)).get(w.name, {}).get(dominant)
            if hw:
                roofline["traffic"] = hw["dram_bytes"]
                sm_hz = (clocks or {}).get("sm_mhz") or 1965.0
                issue_peak = 148 * 4 * sm_hz * 1e6  # warp instructions / s: 148 SMs x 4 schedulers x SM clock
                roofline["issue"] = {"warp_inst_per_launch": hw["warp_inst"], "active_lanes_per_inst": hw["lanes_per_inst"],
                                     "achieved_ginst_s": hw["warp_inst"] / (dom_ms * 1e-3) / 1e9,
                                     "peak_ginst_s": issue_peak / 1e9,
                                     "frac": hw["warp_inst"] / (dom_ms * 1e-3) / issue_peak,
                                     "source": "ncu counters of the same kernel (profiles/traffic.json) over the live launch time"}
        except Exception:
            pass
    if roofline["frac"] > 1.0:
        roofline["note"] += ("; frac > 1 is expected here: the algorithmic bytes are PER-RAY fetches of the reference "
                             "traversal, and the 32 coherent rays of a warp share one fetch through L1/L2 -- DRAM traffic "
                             "(`traffic`) is <1 % of it and the kernel is bound by instruction issue (`issue`), see DESIGN.md")

    # ---- CPU baseline + parity spot check (outside every timed region) ---------------------------------------------
    cpu_baseline, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        cpu = CpuPath(w, mesh_files, hs.arrays)
        best = None
        for _ in range(2):
            dt, r, t_cpu, vis_cpu = cpu.frame()
            best = dt if best is None else min(best, dt)
        cpu_baseline = cpu.describe(best, r)
        gpu_vis = d_vis[0].cpu().numpy() if nl else np.zeros(w.pixels, np.uint8)
        t_gpu = np.where(res_hits["prim"] != capi.MISS, res_hits["t"], np.float32(np.inf)).astype(np.float32)
        parity = {"against": cpu.kind, "t_bit_mismatches": int((t_gpu.view(np.uint32) != t_cpu.view(np.uint32)).sum()),
                  "visibility_mismatches": int((gpu_vis != vis_cpu).sum()) if nl else 0, "rays": int(r)}

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

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": w.name, "description": w.description, "mesh": workloads.mesh_label(w),
                       "width": w.width, "height": w.height, "primary_rays": w.pixels, "shadow_rays": shadow_rays * nl,
                       "triangles": sizes.num_triangles, "kd_nodes": sizes.num_nodes, "tri_lanes": sizes.num_lanes,
                       "tile": args.tile, "parallelism": f"image tiles round-robin over {world} GPU(s), scene replicated"
                                      + ("; hit-record gather to rank 0 (NCCL) overlapped with the shadow pass, "
                                         "frame re-assembled by dodrt_frame_assemble_device" if world > 1 else ""),
                       "l2": "flushed between steps (512 MiB memset outside the per-step event bracket)",
                       "host_build_s": round(build_s, 2), "upload_s": round(upload_s, 2), "wall_s_timed_region": round(wall, 3)},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "parity": parity, "frame_ms": ms_per_step, "reference_frame": reference_frame}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
