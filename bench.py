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
  configs (N=1): the other BASELINE configs, each measured in this run with its own context — C1 Cornell 512^2 depth 4,
           C3 1000 instances x 100 k triangles + 1000 spheres, C4 64 two-key motion instances + spheres + 10 k curve segments:
           Mrays/s, ms per subframe, stage times, rays by type
  c5 (every N): BASELINE configs[4] — the C2 scene at 3840x2160, 128 subframes of 8 spl IN TOTAL split over the ranks (strong
           scaling), render and reduce timed separately
  roofline_issue: the bound the extend kernel is actually against — warp instructions issued per second vs SMs x 4 issue
           slots x SM clock; instructions per ray come from the newest committed ncu capture, rays/s and clock from this run
  traversal_counters: wide nodes and primitive tests per ray of one subframe, counted on the device in this run by the
           -DRT3_STATS build of the same sources (rendertoy3c_b200/librt3_stats.so)
  N>1 adds reduce_check (N subframes sample-partitioned + all-reduce vs the same N subframes on rank 0 alone) and cpp_host (the
           single-process C++ host with --gpus N, i.e. rt3_allreduce_accum over NCCL)
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
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C4 / C5 records of the N=1 line")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 record (3840x2160, --c5-subframes subframes in total over all ranks)")
    ap.add_argument("--c5-subframes", type=int, default=128, help="C5: subframes of 8 spl in total over all ranks (128 = 1024 spp)")
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
    OUT.emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tot_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": dict(config_dict(args, args.gpus), film_rendered_per_step="%dx%d" % (w, h), paths_rendered_per_step=w * h * SPL,
                                             note="same scene, camera, depth and samples per pixel; the film is reduced so that a CPU step stays bounded (Mrays/s is per ray)"),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "samples_per_s": tot_samples / tot_s,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is OptiX/RT-core only (no CPU renderer, not buildable offline); this arm is the in-repo scalar C++ port on all host cores",
    }))


# --------------------------------------------------------------------------------------------- GPU arm
class _CudaArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


EXTEND_KERNELS = ("k_traverse<0", "k_extend_packets")   # the kernels of the extend stage, as the ncu launch list names them


def issue_evidence():
    """Issue-slot utilisation of the extend launches from the newest committed full ncu capture of this command
    (profiles/r*_ncu_bench_traverse_full.txt): static evidence, not measured in this run."""
    import glob
    import re
    cand = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_bench_traverse_full.txt")))
    if not cand:
        return None
    txt = open(cand[-1]).read()
    blocks = [b for b in txt.split("=" * 100) if any(k in b for k in EXTEND_KERNELS)]
    vals = {}
    for key in ("smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
                "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"):
        v = [float(m.group(1)) for b in blocks for m in [re.search(re.escape(key) + r"\s+([0-9.]+)", b)] if m]
        if v:
            vals[key] = round(sum(v) / len(v), 2)
    return {"source": "profiles/" + os.path.basename(cand[-1]), "extend_launches_averaged": len(blocks), **vals}


def issue_roofline(kernel_prefix, ext_ms_per_step, sm_mhz):
    """The bound the traversal kernel is actually against: warp instructions issued per second vs SMs x 4 schedulers x
    SM clock.  Instruction counts need ncu: they come from the newest committed launch list of this command that carries
    smsp__inst_executed.sum (profiles/r*_launch_shares.json); the kernel time and the clock are this run's."""
    import glob
    sms, per_sm = 148, 4
    clock_hz = (sm_mhz or 1965.0) * 1e6
    peak = sms * per_sm * clock_hz
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_launch_shares.json")), reverse=True):
        ks = json.load(open(path))["kernels"]
        inst = sum(v.get("warp_inst", 0) for k, v in ks.items() if k.startswith(kernel_prefix))   # kernel_prefix: one prefix or a tuple of them
        if inst > 0 and ext_ms_per_step > 0:
            ach = inst / (ext_ms_per_step * 1e-3)
            return {"bound": "issue", "achieved": ach / 1e9, "peak": peak / 1e9, "unit": "G warp-inst/s", "frac": ach / peak,
                    "warp_instructions_per_step": inst, "instructions_source": "profiles/" + os.path.basename(path) + " (ncu smsp__inst_executed.sum, same command)",
                    "kernel_ms_per_step": ext_ms_per_step, "sm_mhz": sm_mhz, "peak_is": "148 SMs x 4 issue slots x SM clock"}
    ev = issue_evidence()
    if not ev or "smsp__issue_active.avg.pct_of_peak_sustained_active" not in ev:
        return None
    frac = ev["smsp__issue_active.avg.pct_of_peak_sustained_active"] / 100.0
    return {"bound": "issue", "achieved": frac * peak / 1e9, "peak": peak / 1e9, "unit": "G warp-inst/s", "frac": frac, "source": ev["source"],
            "note": "no committed launch list with instruction counts yet: fraction = issue-slot utilisation of the committed full capture", "peak_is": "148 SMs x 4 issue slots x SM clock"}


def time_subframes(g, make, warmup, steps, stream, torch):
    """device-timed render of `steps` subframes after `warmup`: (ms, stats of the timed region)"""
    for i in range(warmup):
        g.launch_subframe(make(i))
    g.sync()
    g.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for i in range(steps):
        g.launch_subframe(make(warmup + i))
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), g.stats()


def run_config(name, desc, local, torch, steps=3, warmup=2, spl=SPL, max_depth=None, options=None, counters=True):
    """one BASELINE config measured like the headline: own context, device-resident, CUDA events on the library's stream"""
    from rendertoy3c_b200 import scenes
    from rendertoy3c_b200.api import Context, make_settings
    t0 = time.perf_counter()
    g = Context(local)
    for k, v in (options or {}).items():
        g.set_option(k, v)
    scenes.replay(desc, g)
    g.sync()
    build_s = time.perf_counter() - t0
    uvw = g.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, desc.width / desc.height)
    stream = torch.cuda.ExternalStream(g.stream(), device=torch.device("cuda", local))
    depth = desc.max_depth if max_depth is None else max_depth

    def make(i):
        return make_settings(desc, uvw, i, samples_per_launch=spl, max_depth=depth)

    ms, st = time_subframes(g, make, warmup, steps, stream, torch)
    rays = st["rays_primary"] + st["rays_bounce"] + st["rays_shadow"]
    assert st["error_flags"] == 0, name + ": traversal stack overflow"
    g.set_option("timing", 1)
    g.reset_stats()
    g.launch_subframe(make(warmup + steps))
    g.sync()
    s = g.stats()
    g.set_option("timing", 0)
    out = {"workload": desc.name, "instanced_primitives": int(desc.total_instanced_prims()), "width": desc.width, "height": desc.height, "samples_per_step": spl,
           "max_depth": depth, "steps": steps, "warmup": warmup, "value": rays / (ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms / steps,
           "samples_per_s": st["samples"] / (ms * 1e-3), "rays_per_step": {k: st["rays_" + k] // steps for k in ("primary", "bounce", "shadow")},
           "stage_ms": {k: s[k] for k in ("ms_generate", "ms_extend", "ms_shade", "ms_connect", "ms_resolve", "ms_total")},
           "upload_and_build_s": build_s, "flattened_instances": s["flattened_instances"], "traversal_passes": s["traversal_passes"]}
    g.close()
    if options:
        out["options"] = options
    if counters:
        out["traversal_counters"] = traversal_counters(desc, local)
    return out


def traversal_counters(desc, local):
    """wide nodes visited and primitive tests per ray, counted on the device by the -DRT3_STATS build of the same sources
    (one subframe of at most 960x540 x 8 spl: the counters are 32 bit)"""
    from rendertoy3c_b200 import scenes
    from rendertoy3c_b200.api import Context, make_settings
    lib = os.path.join(ROOT, "rendertoy3c_b200", "librt3_stats.so")
    if not os.path.exists(lib):
        return {"unavailable": "rendertoy3c_b200/librt3_stats.so not built"}
    g = Context(local, lib_path=lib)
    scenes.replay(desc, g)
    w, h = (960, 540) if desc.width * desc.height > 960 * 540 else (desc.width, desc.height)
    uvw = g.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, w / h)
    g.debug_counters()
    g.reset_stats()
    g.launch_subframe(make_settings(desc, uvw, 0, samples_per_launch=SPL, width=w, height=h, max_depth=desc.max_depth))
    g.sync()
    c = g.debug_counters()
    st = g.stats()
    g.close()
    rays = max(1, int(c[5]) // max(1, st["traversal_passes"]))   # a split traversal finishes every ray twice (pass 1, pass 2)
    return {"wide_nodes_per_ray": c[2] / rays, "primitive_tests_per_ray": c[3] / rays, "rounds_per_ray": c[4] / rays, "instance_entries_per_ray": c[13] / rays, "rays_counted": rays, "traversal_passes": st["traversal_passes"],
            "rays_by_type": {k: st["rays_" + k] for k in ("primary", "bounce", "shadow")}, "film": "%dx%d x %d spl, all ray types of one subframe" % (w, h, SPL),
            "library": "rendertoy3c_b200/librt3_stats.so (-DRT3_STATS build of the same sources; not the timed library)"}


def cpp_host_run(n_gpus, local_root):
    """the single-process C++ host (rendertoy3c_b200/host/wavefront.cpp) with --gpus N: loadOBJ, one context per GPU,
    rt3_allreduce_accum (NCCL resolved with dlopen) — run once from rank 0 after the timed work"""
    import tempfile
    from rendertoy3c_b200 import scenes
    exe = os.path.join(ROOT, "rendertoy3c_b200", "host", "wavefront")
    if not os.path.exists(exe):
        return {"unavailable": "rendertoy3c_b200/host/wavefront not built"}
    d = tempfile.mkdtemp(prefix="rt3_host_")
    desc = scenes.cornell(width=256, height=256)
    obj = os.path.join(d, "cornell.obj")
    scenes.write_obj(desc, obj)
    c = desc.camera
    res = {}
    frames = {}
    for n in (1, n_gpus):
        out = os.path.join(d, "out_%d.ppm" % n)
        cmd = [exe, "--scene", obj, "--width", "256", "--height", "256", "--spp", str(8 * max(8, n_gpus)), "--max-depth", "4", "--fovy", repr(c.fovy), "--gpus", str(n), "--out", out,
               "--eye", *map(repr, c.eye), "--lookat", *map(repr, c.lookat), "--up", *map(repr, c.up)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            return {"failed": (r.stderr or r.stdout)[-400:], "gpus": n}
        res["gpus_%d" % n] = json.loads(r.stdout.strip().splitlines()[-1])
        with open(out, "rb") as f:
            frames[n] = f.read()
    import numpy as np
    a = np.frombuffer(frames[1][-256 * 256 * 3:], dtype=np.uint8).astype(int)
    b = np.frombuffer(frames[n_gpus][-256 * 256 * 3:], dtype=np.uint8).astype(int)
    res["frame_max_lsb_vs_1gpu"] = int(np.abs(a - b).max())
    res["what"] = "Cornell 256x256, %d spp through rt3_allreduce_accum on %d GPUs vs the same render on one GPU" % (8 * max(8, n_gpus), n_gpus)
    return res


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
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, Wm = args.steps, args.warmup
    desc = scenes.terrain(n=args.grid, width=args.width, height=args.height)
    g = Context(local)
    scenes.replay(desc, g)
    uvw = g.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, args.width / args.height)
    stream = torch.cuda.ExternalStream(g.stream(), device=dev)
    accum_mode = 1 if world > 1 else 0

    def settings(i):  # rank r renders subframes r, r+N, ...
        return make_settings(desc, uvw, rank + world * i, samples_per_launch=SPL, accum_mode=accum_mode, max_depth=8)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_and_finalize(ctx, nsub_total):
        if world > 1:
            ptr, n = ctx.accum_device_ptr()
            t = torch.as_tensor(_CudaArray(ptr, n), device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            ctx.finalize_accum(nsub_total)

    # ---- warm-up
    for i in range(Wm):
        g.launch_subframe(settings(i))
    g.sync()
    if world > 1:
        g.clear_accum()
        reduce_and_finalize(g, world)
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
        reduce_and_finalize(g, K * world)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    st = g.stats()
    launches = st["kernel_launches"] - l0
    rays = st["rays_primary"] + st["rays_bounce"] + st["rays_shadow"]
    rays_by_type = {k: st["rays_" + k] for k in ("primary", "bounce", "shadow")}
    samples = st["samples"]
    assert st["error_flags"] == 0, "traversal stack overflow"

    # ---- e2e: per step settings H2D (inside rt3_launch_subframe as kernel parameters) + frame D2H into pinned memory
    frames = [torch.empty((args.height, args.width, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
    frame = frames[0]
    g.clear_accum()
    g.reset_stats()
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        g.launch_subframe(settings(Wm + i))
        if world == 1:
            g.download_frame_async_into(frames[i & 1].data_ptr())   # the copy runs beside the next subframe; g.sync() below completes the last one
    if world > 1:
        g.sync()
        reduce_and_finalize(g, K * world)
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

    # ---- N > 1: the partitioned image equals the single-GPU image (fp32 summation order apart)
    reduce_check = None
    if world > 1:
        g.clear_accum()
        g.launch_subframe(make_settings(desc, uvw, rank, samples_per_launch=SPL, accum_mode=1, max_depth=8))     # subframe index = rank
        g.sync()
        reduce_and_finalize(g, world)
        if rank == 0:
            part_accum, part_frame = g.download_accum(), g.download_frame()
            g.clear_accum()
            for sf in range(world):
                g.launch_subframe(make_settings(desc, uvw, sf, samples_per_launch=SPL, accum_mode=1, max_depth=8))
            g.sync()
            g.finalize_accum(world)
            one_accum, one_frame = g.download_accum(), g.download_frame()
            a, b = part_accum[..., :3].astype(np.float64), one_accum[..., :3].astype(np.float64)
            reduce_check = {"max_rel_diff": float((np.abs(a - b) / np.maximum(np.abs(b), 1e-3)).max()), "frame_max_lsb": int(np.abs(part_frame.astype(int) - one_frame.astype(int)).max()),
                            "bit_identical_pixels": float((part_accum.view(np.uint32) == one_accum.view(np.uint32)).all(axis=-1).mean()),
                            "what": "%d subframes x %d spl at %dx%d: one per rank + NCCL sum + finalize, vs the same %d subframes on rank 0 alone" % (world, SPL, args.width, args.height, world)}
        barrier()
    g.close()

    # ---- C5 as BASELINE names it: 3840x2160, 1024 spp = 128 subframes of 8 IN TOTAL, split over the ranks (strong scaling)
    c5 = None
    if not args.no_c5:
        d5 = scenes.terrain(n=args.grid, width=3840, height=2160)
        g5 = Context(local)
        scenes.replay(d5, g5)
        uvw5 = g5.camera_uvw(d5.camera.eye, d5.camera.lookat, d5.camera.up, d5.camera.fovy, 3840 / 2160)
        stream5 = torch.cuda.ExternalStream(g5.stream(), device=dev)
        total_sub = args.c5_subframes
        mine = [sf for sf in range(total_sub) if sf % world == rank]
        for sf in mine[:1]:
            g5.launch_subframe(make_settings(d5, uvw5, sf, samples_per_launch=SPL, accum_mode=1, max_depth=8))
        g5.sync()
        if world > 1:
            reduce_and_finalize(g5, world)
        g5.clear_accum()
        g5.reset_stats()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        barrier()
        ev[0].record(stream5)
        for sf in mine:
            g5.launch_subframe(make_settings(d5, uvw5, sf, samples_per_launch=SPL, accum_mode=1, max_depth=8))
        ev[1].record(stream5)
        g5.sync()
        if world > 1:
            ptr, n = g5.accum_device_ptr()
            t = torch.as_tensor(_CudaArray(ptr, n), device=dev)
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()   # the ranks finish rendering at slightly different times: that skew is already in max(render_ms), not reduce time
            r0.record()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            r1.record()
            torch.cuda.synchronize()
            reduce_ms = r0.elapsed_time(r1)
        else:
            reduce_ms = 0.0
        g5.finalize_accum(total_sub)
        ev[2].record(stream5)
        barrier()
        render_ms = ev[0].elapsed_time(ev[1])
        s5 = g5.stats()
        r5 = s5["rays_primary"] + s5["rays_bounce"] + s5["rays_shadow"]
        vals = torch.tensor([render_ms, reduce_ms], dtype=torch.float64, device="cuda")
        cnt = torch.tensor([r5, s5["samples"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(vals, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        total_ms = float(vals[0]) + float(vals[1])
        c5 = {"workload": "C5: C2 scene at 3840x2160, %d subframes x %d spl = %d spp in total, split over %d GPU(s) (strong scaling)" % (total_sub, SPL, total_sub * SPL, world),
              "subframes_total": total_sub, "subframes_per_gpu": len(mine), "render_ms_max_over_ranks": float(vals[0]), "reduce_ms_max_over_ranks": float(vals[1]),
              "reduce_bytes": 3840 * 2160 * 16, "total_ms": total_ms, "value": float(cnt[0]) / (total_ms * 1e-3) / 1e6, "unit": UNIT,
              "samples_per_s": float(cnt[1]) / (total_ms * 1e-3), "ms_per_subframe_per_gpu": float(vals[0]) / max(1, len(mine))}
        g5.close()

    # ---- max over ranks / totals
    if world > 1:
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
        c = torch.tensor([rays, samples, rays2, launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        rays, samples, rays2, launches = (int(x) for x in c.tolist())
        dist.barrier()
        dist.destroy_process_group()
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
            ks = [v for n, v in json.load(open(shares))["kernels"].items() if n.startswith(EXTEND_KERNELS)]
            if ks:
                k = {f: sum(v[f] for v in ks) for f in ("dram_read_MB", "dram_write_MB", "launches")}
                traffic = {"bytes_per_step": (k["dram_read_MB"] + k["dram_write_MB"]) * 1e6, "launches_per_step": k["launches"],
                           "algorithmic_bytes_per_step": BYTES_PER_RAY_EXTEND * ext_rays / 2, "source": "profiles/%s (ncu dram__bytes_read/write.sum of this command)" % os.path.basename(shares)}
        out = {
            "metric": METRIC, "value": rays / (ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, world),
            "samples_per_s": samples / (ms * 1e-3),
            "rays_per_step": rays / K / world,
            "rays_by_type_rank0": rays_by_type,
            "e2e": {"value": rays2 / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": C.sizeof(RenderSettings),
                    "d2h_bytes_per_step": args.width * args.height * 4 if world == 1 else args.width * args.height * 4 // K,
                    "what": "rt3_launch_subframe(host settings) + rt3_download_frame_async(pinned host u8 frame, two buffers in turn) per step, rt3_sync at the end, wall clock"},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "the extend stage: rt3::k_extend_packets (camera rays, depth 0) + rt3::k_traverse<0> (bounce rays), closest hit", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_ray": BYTES_PER_RAY_EXTEND,
                         "kernel_ms_per_step": ext_ms / 2, "kernel_Mrays_s": ext_rays / (ext_ms * 1e-3) / 1e6,
                         "kernel_share_of_step": ext_ms / tot_ms,
                         "connect_kernel_Mrays_s": con_rays / (con_ms * 1e-3) / 1e6 if con_ms > 0 else None,
                         "whole_path_achieved_GBs": BYTES_PER_RAY_PATH * rays / (ms * 1e-3) / 1e9,
                         "note": "HBM is the roofline of the queue traffic only: the kernel reads and writes its algorithmic bytes once (traffic ~ 1.0x) and is "
                                 "bound by instruction issue, see roofline_issue",
                         "issue_bound_evidence": issue_evidence()},
            "roofline_issue": issue_roofline(EXTEND_KERNELS, ext_ms / 2, (clk or {}).get("sm_mhz")),
            "stage_ms_last_step": stage,
        }
        if reduce_check is not None:
            out["reduce_check"] = reduce_check
        if c5 is not None:
            out["c5"] = c5
        if world == 1:
            out["traversal_counters"] = traversal_counters(desc, local)
            if not args.no_configs:
                cfg = {}
                cfg["C1"] = run_config("C1", scenes.cornell(width=512, height=512), local, torch, steps=4, warmup=2)
                cfg["C3"] = run_config("C3", scenes.instanced(width=1920, height=1080), local, torch)
                # the same scene with its 1000 instances kept as instances (one BLAS, 1000 TLAS leaves): the two-level path proper
                cfg["C3_instances_kept"] = run_config("C3", scenes.instanced(width=1920, height=1080), local, torch, options={"flatten": 0}, counters=False)
                cfg["C4"] = run_config("C4", scenes.motion(width=1920, height=1080), local, torch)
                out["configs"] = cfg
            if not args.no_cpu_baseline:
                g2 = Context(local)
                scenes.replay(desc, g2)
                out["cpu_baseline"] = cpu_baseline(desc, gpu=g2)
                g2.close()
        else:
            out["cpp_host"] = cpp_host_run(world, ROOT)
        OUT.emit(json.dumps(out))


class QuietStdout:
    """Everything libraries print to fd 1 while the bench runs (NCCL's version banner, ...) goes to stderr, so that stdout carries
    exactly ONE line: the JSON record printed through emit()."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(text, flush=True)
        os.dup2(2, 1)


OUT = None


if __name__ == "__main__":
    a = parse()
    OUT = QuietStdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_rt3(a)
