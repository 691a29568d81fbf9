#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json: Mrays/s, primary + bounce + shadow).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (N=1): BASELINE.json configs[1] — "C2": synthetic tessellated textured mesh, 1,002,528
triangles (+2 emissive), single BLAS, 1920x1080, depth 8, Lambert + uniform-light NEE, faithful
estimator.  A STEP is one subframe = samples_per_launch (8) paths per pixel = 16.6 M paths through
generate/extend/shade/connect/resolve (K=8 steps = the config's 64 spp).
N>1: scene replicated, rank r renders subframe indices r, r+N, ... (weak scaling: K subframes per
GPU); one NCCL all-reduce (sum) of the float4 accumulation buffers + a fused normalise/quantise
closes the job and is inside the timed region.
  value  : whole-job Mrays/s, inputs resident in HBM, CUDA events on the library's stream
  e2e    : same metric through the C ABI per step with HOST buffers: settings struct in (H2D),
           8-bit frame out (D2H into pinned memory) every step
  roofline: dominant kernel k_traverse<extend>: algorithmic queue bytes (48 B ray read + 20 B hit
           write = 68 B/ray) / its CUDA-event time vs the measured HBM copy peak
  cpu_baseline: the scalar C++ oracle (oracle/, "port": the reference has no CPU renderer and its
           OptiX path cannot be built here) on all host cores, on a bounded sample of the same scene
--impl reference: that CPU oracle as the reference arm (the reference itself is OptiX-only).
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

METRIC = "Mrays/s (primary+bounce+shadow)"
UNIT = "Mrays/s"
SPL = 8
WORKLOAD = "C2 terrain 1,002,530 tris single BLAS 1920x1080 depth 8 NEE faithful"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rt3")
    ap.add_argument("--grid", type=int, default=708, help="terrain quads per side (708 -> 1,002,528 tris)")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def config_dict(args, n_gpus):
    return {"workload": WORKLOAD if (args.grid, args.width, args.height) == (708, 1920, 1080) else
            "terrain grid=%d %dx%d depth 8" % (args.grid, args.width, args.height),
            "triangles": 2 * args.grid * args.grid + 2, "width": args.width, "height": args.height,
            "samples_per_step": SPL, "paths_per_step": args.width * args.height * SPL, "max_depth": 8,
            "parallelism": "sample-partitioned x%d (scene replicated)" % n_gpus,
            "l2_policy": "inputs larger than L2: ~3.3 GB of queue planes per step vs 126 MB L2"}


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU arm
def oracle_scene(desc):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_backend import OracleScene
    from rendertoy3c_b200 import scenes
    o = OracleScene(nthreads=0)
    scenes.replay(desc, o)
    return o


def cpu_sample(o, desc, width, height, subframe, seconds_hint=None):
    """one bounded oracle step: a (width x height) x 8-spl subframe of the same scene/camera"""
    from rendertoy3c_b200.api import make_settings
    uvw = o.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, width / height)
    o.reset_stats()
    t = time.perf_counter()
    o.launch_subframe(make_settings(desc, uvw, subframe, samples_per_launch=SPL, width=width, height=height, max_depth=8))
    dt = time.perf_counter() - t
    st = o.stats()
    rays = st["rays_primary"] + st["rays_bounce"] + st["rays_shadow"]
    return rays, dt, st["samples"]


def cpu_baseline(desc, target_s=12.0, gpu=None):
    """Times the oracle port on a bounded sample; with `gpu` (the product context, after its timed region) the same
    subframe is also rendered on the GPU and compared: the checker use of the oracle (same-seed relMSE)."""
    o = oracle_scene(desc)
    cores = os.cpu_count() or 1
    rays, dt, _ = cpu_sample(o, desc, 96, 54, 0)  # probe
    rate = rays / dt
    # size the sample for ~target_s of CPU work, same aspect ratio
    scale = max(1.0, min(20.0, (target_s * rate / rays) ** 0.5))
    w, h = int(96 * scale) // 8 * 8, int(54 * scale) // 2 * 2
    rays, dt, samples = cpu_sample(o, desc, w, h, 0)
    out = {"value": rays / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "%dx%d px x %d spl subframe of the same scene and camera (%d rays, %.1f s, BVH2 build excluded)" % (w, h, SPL, rays, dt),
           "samples_per_s": samples / dt}
    if gpu is not None:
        import numpy as np
        from rendertoy3c_b200.api import make_settings
        a = o.download_accum()[..., :3].astype(np.float64)
        uvw = gpu.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, w / h)
        gpu.launch_subframe(make_settings(desc, uvw, 0, samples_per_launch=SPL, width=w, height=h, max_depth=8))
        b = gpu.download_accum()[..., :3].astype(np.float64)
        out["same_seed_check"] = {"relMSE_gpu_vs_cpu": float(np.mean((b - a) ** 2 / (a ** 2 + 1e-2))), "max_abs_diff": float(np.abs(b - a).max()),
                                  "bit_identical": bool(np.array_equal(o.download_accum().view(np.uint32), gpu.download_accum().view(np.uint32))),
                                  "what": "float accumulation buffer of that subframe, GPU (librt3.so) vs CPU oracle"}
    o.close()
    return out


def run_reference(args):
    """--impl reference: the reference's path on the host cores.  The reference has no CPU
    implementation and its OptiX path cannot be compiled here, so this arm times the
    oracle port of the same path (all host threads) on bounded samples of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rendertoy3c_b200 import scenes
    desc = scenes.terrain(n=args.grid, width=args.width, height=args.height)
    o = oracle_scene(desc)
    cores = os.cpu_count() or 1
    rays, dt, _ = cpu_sample(o, desc, 96, 54, 0)
    per_step_s = max(1.0, min(15.0, 120.0 / max(1, args.steps + args.warmup)))
    scale = max(1.0, min(20.0, (per_step_s * (rays / dt) / rays) ** 0.5))
    w, h = int(96 * scale) // 8 * 8, int(54 * scale) // 2 * 2
    for i in range(args.warmup):
        cpu_sample(o, desc, w, h, i)
    tot_rays, tot_s, tot_samples = 0, 0.0, 0
    for i in range(args.steps):
        r, s, sm = cpu_sample(o, desc, w, h, args.warmup + i)
        tot_rays += r
        tot_s += s
        tot_samples += sm
    v = tot_rays / tot_s / 1e6
    sample = "%dx%d px x %d spl subframe per step (same scene, camera, depth)" % (w, h, SPL)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tot_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "samples_per_s": tot_samples / tot_s,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is OptiX/RT-core only (no CPU renderer, not buildable offline); this arm is the in-repo scalar C++ port on all host cores",
    }))


# --------------------------------------------------------------------------------------------- GPU arm
class _CudaArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def issue_evidence():
    """Issue-slot utilisation of the extend launches from the newest committed full ncu capture of this command
    (profiles/r*_ncu_bench_traverse_full.txt): static evidence, not measured in this run."""
    import glob
    import re
    cand = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_bench_traverse_full.txt")))
    if not cand:
        return None
    txt = open(cand[-1]).read()
    blocks = [b for b in txt.split("=" * 100) if "k_traverse<0" in b]
    vals = {}
    for key in ("smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
                "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"):
        v = [float(m.group(1)) for b in blocks for m in [re.search(re.escape(key) + r"\s+([0-9.]+)", b)] if m]
        if v:
            vals[key] = round(sum(v) / len(v), 2)
    return {"source": "profiles/" + os.path.basename(cand[-1]), "extend_launches_averaged": len(blocks), **vals}


def run_rt3(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from rendertoy3c_b200 import scenes
    from rendertoy3c_b200._abi import RenderSettings
    from rendertoy3c_b200.api import Context, make_settings

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; librt3 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, Wm = args.steps, args.warmup
    desc = scenes.terrain(n=args.grid, width=args.width, height=args.height)
    g = Context(local)
    scenes.replay(desc, g)
    uvw = g.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, args.width / args.height)
    stream = torch.cuda.ExternalStream(g.stream(), device=torch.device("cuda", local))
    accum_mode = 1 if world > 1 else 0

    def settings(i):  # rank r renders subframes r, r+N, ...
        return make_settings(desc, uvw, rank + world * i, samples_per_launch=SPL, accum_mode=accum_mode, max_depth=8)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_and_finalize(nsub):
        if world > 1:
            ptr, n = g.accum_device_ptr()
            t = torch.as_tensor(_CudaArray(ptr, n), device=torch.device("cuda", local))
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            g.finalize_accum(nsub * world)

    # ---- warm-up
    for i in range(Wm):
        g.launch_subframe(settings(i))
    g.sync()
    if world > 1:
        g.clear_accum()
        reduce_and_finalize(1)
        g.clear_accum()
    g.sync()

    # ---- timed: device-resident
    g.reset_stats()
    l0 = g.stats()["kernel_launches"]
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(K):
        g.launch_subframe(settings(Wm + i))
    if world > 1:
        g.sync()
        reduce_and_finalize(K)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    st = g.stats()
    launches = st["kernel_launches"] - l0
    rays = st["rays_primary"] + st["rays_bounce"] + st["rays_shadow"]
    samples = st["samples"]
    assert st["error_flags"] == 0, "traversal stack overflow"

    # ---- e2e: per step settings H2D (inside rt3_launch_subframe as kernel parameters) + frame D2H into pinned memory
    frame = torch.empty((args.height, args.width, 4), dtype=torch.uint8).pin_memory()
    g.clear_accum()
    g.reset_stats()
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        g.launch_subframe(settings(Wm + i))
        if world == 1:
            g.download_frame_into(frame.data_ptr())
    if world > 1:
        g.sync()
        reduce_and_finalize(K)
        g.download_frame_into(frame.data_ptr())
    g.sync()
    barrier()
    e2e_s = time.perf_counter() - t0
    st2 = g.stats()
    rays2 = st2["rays_primary"] + st2["rays_bounce"] + st2["rays_shadow"]
    clk = clocks.stop() if rank == 0 else None

    # ---- per-kernel time of the dominant kernel (CUDA events on the library's stream, per stage)
    g.set_option("timing", 1)
    g.reset_stats()
    ext_ms, con_ms, ext_rays, con_rays, tot_ms = 0.0, 0.0, 0, 0, 0.0
    for i in range(2):
        g.reset_stats()
        g.launch_subframe(settings(Wm + i))
        g.sync()
        s = g.stats()
        ext_ms += s["ms_extend"]; con_ms += s["ms_connect"]; tot_ms += s["ms_total"]
        ext_rays += s["rays_primary"] + s["rays_bounce"]; con_rays += s["rays_shadow"]
        stage = {k: s[k] for k in ("ms_generate", "ms_extend", "ms_shade", "ms_connect", "ms_resolve", "ms_total")}
    g.set_option("timing", 0)

    # ---- max over ranks / totals
    if world > 1:
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
        c = torch.tensor([rays, samples, rays2, launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        rays, samples, rays2, launches = (int(x) for x in c.tolist())
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        BYTES_PER_RAY_EXTEND = 68.0   # 48 B ray record read + 20 B hit record written
        BYTES_PER_RAY_PATH = 240.0    # whole wavefront segment, all stages (SURVEY 8d's figure restated for this layout)
        achieved = BYTES_PER_RAY_EXTEND * ext_rays / (ext_ms * 1e-3) / 1e9
        # DRAM bytes of the extend launches of one step, from the committed ncu capture of this command
        traffic = None
        import glob
        cand = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_launch_shares.json")))  # newest committed capture
        shares = cand[-1] if cand else ""
        if shares and (args.grid, args.width, args.height) == (708, 1920, 1080):
            k = json.load(open(shares))["kernels"].get("k_traverse<0, 1>")
            if k:
                traffic = {"bytes_per_step": (k["dram_read_MB"] + k["dram_write_MB"]) * 1e6, "launches_per_step": k["launches"],
                           "algorithmic_bytes_per_step": BYTES_PER_RAY_EXTEND * ext_rays / 2, "source": "profiles/%s (ncu dram__bytes_read/write.sum of this command)" % os.path.basename(shares)}
        out = {
            "metric": METRIC, "value": rays / (ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, world),
            "samples_per_s": samples / (ms * 1e-3),
            "rays_per_step": rays / K / world,
            "e2e": {"value": rays2 / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": C.sizeof(RenderSettings),
                    "d2h_bytes_per_step": args.width * args.height * 4 if world == 1 else args.width * args.height * 4 // K,
                    "what": "rt3_launch_subframe(host settings) + rt3_download_frame(pinned host u8 frame) per step, wall clock"},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "rt3::k_traverse<0> (extend, closest hit)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_ray": BYTES_PER_RAY_EXTEND,
                         "kernel_ms_per_step": ext_ms / 2, "kernel_Mrays_s": ext_rays / (ext_ms * 1e-3) / 1e6,
                         "kernel_share_of_step": ext_ms / tot_ms,
                         "connect_kernel_Mrays_s": con_rays / (con_ms * 1e-3) / 1e6 if con_ms > 0 else None,
                         "whole_path_achieved_GBs": BYTES_PER_RAY_PATH * rays / (ms * 1e-3) / 1e9,
                         "note": "software BVH traversal is latency/issue bound, not HBM bound (SURVEY 8d): frac is expected to be small; "
                                 "see profiles/ for issue-slot and L2 counters",
                         "issue_bound_evidence": issue_evidence()},
            "stage_ms_last_step": stage,
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(desc, gpu=g)
        print(json.dumps(out))
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_rt3(a)
