#!/usr/bin/env python
"""bench.py -- traced+shaded rays/s (fwd+bwd) of the IRON stage-2 hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--hidden 512] [--patch 64]

Workload (BASELINE.json configs[1]): one stage-2 training step on a 64x64 crop (4096 rays) of the 512x512
colocated-flash fixture view, 8x512 SDF MLP (PE L=6, skip@4, softplus beta=100), three 4x256 material MLPs, GGX
shading, L2 image loss + eikonal, backward (incl. the double backward through the SDF net); synthetic
random-init networks (seed 0) and a synthetic target.  One "step" = trace -> shade -> loss -> backward
[-> gradient allreduce when N > 1].  Rays shard across ranks with replicated weights (weak scaling: every rank
traces its own 4096-ray crop); the only collective is the per-step NCCL allreduce of the flat gradient buffer.

Prints ONE JSON line (rank 0).  `value` is device-timed (CUDA events per step, L2 flushed between steps);
`e2e` runs the same step through the public API from pinned HOST buffers with the H2D/D2H copies in the timed
region; `roofline` is the dominant kernel (the persistent tracer's SDF-MLP tiles); `cpu_baseline` is the CPU
oracle (a PyTorch-CPU restatement of the reference path) timed on this box's host cores.
`--impl reference` times that CPU oracle alone (the reference is pure Python and cannot travel to the GPU box;
the oracle is pinned to it by tests/golden/*).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "traced+shaded rays/sec fwd+bwd (8x512 SDF MLP, GGX)"
# DRAM bytes of one mlp_h16_kernel launch by (hidden width, patch edge), from the ncu capture named below (cold cache: the
# fp16 weight copies and the encoded points are read once from HBM, the activations live in L2)
MLP_H16_TRAFFIC = {(512, 64): 8.18e6}
MLP_H16_TRAFFIC_SOURCE = "profiles/r1l_mlp_h16_ncu.md (ncu --set full, launch 0; a profile constant, not measured by this run)"
UNIT = "rays/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region, sampled synchronously on the main thread of rank 0 right after a
    timed step has been enqueued (the GPU is still executing it), i.e. in the gap between two CUDA-event pairs.

    Why not the recipe's `nvidia-smi -lms 50` loop or an NVML polling thread: on these boxes single NVML queries take up to
    40-130 ms (clock query up to 1.7 s during start-up) and stall concurrent kernel launches; both variants put a 40-120 ms
    outlier into one timed step in about half of the runs (measured, profiles/README.md).  A synchronous query between two
    event pairs cannot overlap a timed step."""
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index: int):
        self.rows = []          # (epoch s, sm MHz, max MHz, [reasons])
        self.slow = 0.0
        self.index = index
        self.nv = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.bits = (("hw_slowdown", getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                         ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                         ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                         ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)))
            self.get_reasons = (getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None)
                                or nv.nvmlDeviceGetCurrentClocksThrottleReasons)
            self.nv = nv
            self.how = "pynvml, synchronous, between timed steps"
        except Exception:
            self.how = "nvidia-smi one-shot queries, between timed steps"

    def sample(self):
        t0 = time.perf_counter()
        try:
            if self.nv is not None:
                sm = float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = int(self.get_reasons(self.h))
                self.rows.append((time.time(), sm, self.mx, [nm for nm, b in self.bits if r & b]))
            else:
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
                parts = [x.strip() for x in out.split(",")]
                self.rows.append((time.time(), float(parts[0]), float(parts[1]),
                                  [nm for nm, v in zip(self.NAMES, parts[2:6]) if v.lower().startswith("active")]))
        except Exception:
            pass
        self.slow = max(self.slow, (time.perf_counter() - t0) * 1e3)

    def stop(self, t_begin=None, t_end=None):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        inside = [r for r in self.rows if t_begin is not None and t_begin <= r[0] <= t_end]
        use = inside if inside else self.rows
        if use:
            out = {"sm_mhz": statistics.median([r[1] for r in use]), "sm_max_mhz": max(r[2] for r in use),
                   "reasons": sorted(set(x for r in use for x in r[3])), "samples": len(use),
                   "samples_inside_timed_region": len(inside), "sampler": self.how,
                   "slowest_query_ms": round(self.slow, 2)}
        return out


def shard_mode() -> str:
    """How the ranks' inputs differ (IRONB_BENCH_SHARD): "views" (default) = every rank renders the canonical centre crop of
    ITS OWN view, the fixture camera orbited about the y-axis in steps of 45 degrees ("extra views by rotating W2C
    about the y-axis in 8 steps") -- distinct rays, targets and eikonal samples per GPU with comparable work (the seed-0
    object is close to a sphere; measured on one GPU the other views cost 3.5-4.0 ms against the canonical view's 4.1 ms, so
    rank 0 stays the slowest and the N-GPU step time is the canonical step plus the exchange); "windows" = one
    view, distinct crop windows around the centre (crop_corner), whose work differs by up to +-5 % (patch/16 stride) or 2x
    (tiled: IRONB_BENCH_CROP_STRIDE=<patch>)."""
    m = os.environ.get("IRONB_BENCH_SHARD", "views")
    return m if m in ("views", "windows") else "views"


def view_w2c(rank: int):
    """W2C of rank's view as a [4,4] float64 numpy array: the fixture pose with the WORLD rotated by rank x 45 degrees about y,
    i.e. the camera orbits the object at its elevation and distance (rank 0 = the canonical fixture view)."""
    import numpy as np
    from oracle import iron_oracle as O   # fixture constants only
    W = np.array(O.FIXTURE_W2C, dtype=np.float64).reshape(4, 4)
    if shard_mode() != "views" or rank % 8 == 0:
        return W
    th = 2.0 * np.pi * (rank % 8) / 8.0
    R = np.eye(4)
    R[0, 0], R[0, 2], R[2, 0], R[2, 2] = np.cos(th), np.sin(th), -np.sin(th), np.cos(th)
    return W @ R


def crop_corner(rank: int, patch: int):
    """Per-rank crop of the 512x512 view ("windows" sharding; with "views" every rank takes the centre crop): rank 0 = the canonical centre crop; the other ranks take DISTINCT windows
    shifted by patch/16 pixels around it (parallel.crop_for_rank).  Weak scaling needs the same work per GPU at every N, and the
    fixture object covers only ~145 pixels of the view: at the 64x64 training patch the shifted windows stay inside the
    silhouette like the centre crop (every ray hits), while windows TILED around the centre (IRONB_BENCH_CROP_STRIDE=<patch>)
    straddle the silhouette and cost about twice the centre crop -- that measures load imbalance, not scaling."""
    from iron_b200.parallel import crop_for_rank
    if shard_mode() == "views" and "IRONB_BENCH_CROP_STRIDE" not in os.environ:
        rank = 0
    stride = int(os.environ.get("IRONB_BENCH_CROP_STRIDE", max(patch // 16, 1)))
    return crop_for_rank(rank, patch, stride=stride)


# ------------------------------------------------------------------------------------------ CPU oracle arm
def oracle_setup(hidden: int, patch: int, rank: int = 0, device=None):
    """The oracle's state for one step: seed-0 weights, the rank's crop, the rank's target / eikonal samples (the same
    seeds run_ours uses).  device: None = CPU (the reference arm / cpu_baseline); a CUDA device = the reference path as
    eager PyTorch on the GPU (cuda_eager_baseline)."""
    import torch
    from oracle import iron_oracle as O
    torch.manual_seed(0)
    mats = O.make_material_dict()
    torch.manual_seed(0)
    sdf = O.make_sdf_params(d_hidden=hidden)
    light = torch.tensor(32.0)
    fx = O.OCamera.fixture()
    K, W2C = fx.K, fx.W2C
    target = torch.rand(patch, patch, 3, generator=torch.Generator().manual_seed(11 + rank)) * 0.5
    eik = torch.empty(patch * patch // 2, 3).uniform_(-1.0, 1.0, generator=torch.Generator().manual_seed(12 + rank))
    if device is not None:
        sdf = {k: v.detach().to(device) for k, v in sdf.items()}
        mats = {n: {k: v.detach().to(device) for k, v in d.items()} for n, d in mats.items()}
        light, K, W2C, target, eik = (t.to(device) for t in (light, K, W2C, target, eik))
    for d in [sdf] + list(mats.values()):
        for v in d.values():
            v.requires_grad_(True)
    light.requires_grad_(True)
    cam = O.OCamera(512, 512, K, W2C).crop(patch, patch, crop_corner(rank, patch))
    return O, sdf, mats, light, cam, target, eik


IMAGE_LOSS = ["reference"]     # set from --loss: "reference" = PyramidL2 + SSIM + roughness range (what the reference trains with)


def oracle_step(state):
    """One oracle step; returns (seconds, loss).  On a CUDA state the time is the host wall time around a synchronised
    step (cuda_eager_baseline additionally takes CUDA events)."""
    import torch
    O, sdf, mats, light, cam, target, eik = state
    for d in [sdf] + list(mats.values()):
        for v in d.values():
            v.grad = None
    light.grad = None
    cuda = light.is_cuda
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    if cuda:
        with torch.device(light.device):       # the oracle's factory calls (linspace, zeros ...) follow its inputs
            loss, _ = O.stage2_step(sdf, mats, light, cam, target, eik.clone(), image_loss=IMAGE_LOSS[0])
        torch.cuda.synchronize()
    else:
        loss, _ = O.stage2_step(sdf, mats, light, cam, target, eik.clone(), image_loss=IMAGE_LOSS[0])
    return time.perf_counter() - t0, float(loss)


def cuda_eager_baseline(hidden: int, patch: int, dev, steps: int = 5, warmup: int = 2):
    """SURVEY 8(d) / BASELINE.md section 3, last bullet: the reference path (models/raytracer.py:778-814 ->
    models/fields.py:120-137 -> models/renderer_ggx.py:82-146, restated by the oracle and pinned to the reference by
    tests/golden) executed as EAGER PyTorch on this B200, fp32 with TF32 off, CUDA-event timed, same weights / crop /
    target / eikonal samples as the hand-written path.  This is the competitor on equal hardware."""
    import torch
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        st = oracle_setup(hidden, patch, 0, device=dev)
        for _ in range(warmup):
            oracle_step(st)
        ms, loss = [], None
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _, loss = oracle_step(st)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    t = statistics.median(ms)
    return {"value": patch * patch / (t * 1e-3), "unit": UNIT, "ms_per_step": t, "steps": steps, "warmup": warmup,
            "loss": loss, "rays": patch * patch,
            "what": "the reference path (oracle restatement, torch.autograd incl. the double backward) as eager PyTorch on "
                    "cuda:0, fp32, allow_tf32=False, CUDA events around the step (host syncs of the reference's control "
                    "flow included), same weights / crop / target / eikonal samples"}


def cpu_baseline(hidden: int, patch: int, budget_s: float = 20.0):
    """The CPU oracle on this box's host cores, bounded to ~budget_s of CPU work: the same step on the same
    weights, on the largest centred sub-crop whose estimated time fits."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # calibrate on a 16x16 crop, then pick the sample size
    st = oracle_setup(hidden, 16)
    t16 = oracle_step(st)[0]
    t16 = min(t16, oracle_step(st)[0])
    per_ray = t16 / 256.0
    size = patch
    while size > 16 and per_ray * size * size * 2 > budget_s:
        size //= 2
    st = oracle_setup(hidden, size)
    oracle_step(st)                      # warm-up
    t1, loss = oracle_step(st)
    ts = [t1]
    if sum(ts) * 2 < budget_s:
        ts.append(oracle_step(st)[0])
    t = statistics.median(ts)
    return {"value": size * size / t, "unit": UNIT, "cores": cores, "kind": "port", "loss": loss, "rays": size * size,
            "sample": f"{size}x{size} centre crop ({size * size} rays) of the same view/weights, {len(ts)} timed step(s) "
                      f"after 1 warm-up, torch CPU fp32 {torch.get_num_threads()} threads, {t:.2f} s/step"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is pure Python/PyTorch and
    /root/reference does not exist on the GPU box, so this times oracle/iron_oracle.py (pinned to the reference by
    tests/golden) with all host threads, on the same config."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st = oracle_setup(args.hidden, 16)
    t16 = min(oracle_step(st)[0], oracle_step(st)[0])
    per_ray = t16 / 256.0
    size = args.patch
    total = args.steps + args.warmup
    while size > 16 and per_ray * size * size * total > 200.0:
        size //= 2
    st = oracle_setup(args.hidden, size)
    for _ in range(args.warmup):
        oracle_step(st)
    ts = [oracle_step(st)[0] for _ in range(args.steps)]
    t = sum(ts) / len(ts)
    v = size * size / t
    sample = (f"{size}x{size} centre crop ({size * size} rays/step) of the configs[1] view, oracle port of the reference "
              f"path, torch CPU fp32, {cores} threads")
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": workload_config(args, args.patch),   # the arm's workload; the bounded sample actually timed is in cpu_baseline.sample
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def ggx_microbench(dev, pk):
    """GGX-only HBM roofline: M = 2^24 points, c~U(0,1), alpha~U(0.01,1), kd,ks~U(0,1), dist~U(1.5,2.5)."""
    import torch
    import iron_b200 as ib
    M = 1 << 24
    g = torch.Generator(device=dev).manual_seed(0)
    c = torch.rand(M, 1, device=dev, generator=g) * 0.98 + 0.01
    v = torch.nn.functional.normalize(torch.randn(M, 3, device=dev, generator=g), dim=-1)
    t = torch.nn.functional.normalize(torch.cross(v, torch.randn(M, 3, device=dev, generator=g), dim=-1), dim=-1)
    n = (c * v + torch.sqrt(1 - c * c) * t).requires_grad_(True)
    alpha = (torch.rand(M, 1, device=dev, generator=g) * 0.99 + 0.01).requires_grad_(True)
    kd = torch.rand(M, 3, device=dev, generator=g).requires_grad_(True)
    ks = torch.rand(M, 3, device=dev, generator=g).requires_grad_(True)
    dist = (torch.rand(M, 1, device=dev, generator=g) + 1.5).requires_grad_(True)
    light = torch.tensor(32.0, device=dev, requires_grad=True)
    rend = ib.GGXColocatedRenderer(use_cuda=True)
    up = torch.randn(M, 3, device=dev, generator=g)
    fwd_ms, bwd_ms = [], []
    for i in range(6):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        out = rend(light, dist, n, v, {"diffuse_albedo": kd, "specular_albedo": ks, "specular_roughness": alpha})
        e1.record()
        torch.cuda.synchronize()
        fwd_ms.append(e0.elapsed_time(e1))
        e1b = torch.cuda.Event(enable_timing=True)
        e1b.record()
        torch.autograd.grad(out["rgb"], [light, dist, n, kd, ks, alpha], grad_outputs=up)
        e2.record()
        torch.cuda.synchronize()
        bwd_ms.append(e1b.elapsed_time(e2))
    f, b = min(fwd_ms[1:]), min(bwd_ms[1:])
    peak = pk["hbm_gbs"]
    # inputs are larger than L2 (M * 56 B = 940 MB); the timed region includes torch's output allocation (no copies)
    return {"points": M, "fwd": {"bytes_per_point": 92, "ms": f, "achieved_gbs": M * 92 / f / 1e6, "frac": M * 92 / f / 1e6 / peak},
            "bwd": {"bytes_per_point": 112, "ms": b, "achieved_gbs": M * 112 / b / 1e6, "frac": M * 112 / b / 1e6 / peak,
                    "note": "autograd.grad through the module: includes d_light zero-init and grad bookkeeping"},
            "peak_gbs": peak, "bound": "hbm", "peak_source": pk["source"]}


def workload_config(args, patch):
    return {"workload": f"BASELINE configs[1]: IRON stage-2 step, {patch}x{patch} crop ({patch * patch} rays/GPU) of the 512x512 "
                        f"colocated-flash fixture view, trace+shade+loss+backward",
            "sdf_mlp": f"8x{args.hidden}, PE L=6, skip@4, softplus(100), weight-norm", "material_mlps": "3 x (4x256, ReLU)",
            "rays_per_gpu": patch * patch, "eikonal_points": patch * patch // 2,
            "sharding": "every rank traces/shades its own rays (rank 0: the canonical centre crop, the others the centre crop of their OWN view (camera orbited about y in 45-degree steps, SURVEY 8d) = distinct rays, comparable work per GPU (3.5-4.1 ms, the canonical view is the most expensive); rank_compute_ms_per_step shows what imbalance is left; IRONB_BENCH_SHARD=windows: distinct crop windows of one view; IRONB_BENCH_CROP_STRIDE=<patch> tiles them), own target/eikonal seeds; gradients packed into one flat buffer by the graph, one in-place all-reduce", "parallelism": f"dp{args.gpus} (rays sharded, weights replicated)",
            "l2": "256 MiB flush between steps, outside the per-step CUDA-event pairs",
            "init": "seed-0 geometric init, light=32",
            "loss": ("PyramidL2 + 1.0 * SSIM(masked) + 0.1 * roughness range + 0.1 * eikonal: the reference's training loss, "
                     "render_surface.py:594-639" if getattr(args, "loss", "reference") == "reference"
                     else "plain L2 on the patch + 0.1 * eikonal"),
            "fill_holes_and_edge_sampling": bool(getattr(args, "driver_defaults", False)),
            "execution": ("one CUDA-graph replay per step (iron_b200.GraphedStage2Step)" if getattr(args, "exec_mode", "graph") == "graph"
                          and getattr(args, "shading", "dense") == "dense" and not getattr(args, "driver_defaults", False)
                          else "eager (Python enqueues every kernel)"),
            "shading": ("dense: every ray of the patch is shaded and non-hit pixels are masked afterwards (no hit-count read-back)"
                        if getattr(args, "shading", "dense") == "dense" else "compact: hits compacted first (one host sync per step)")}


# ------------------------------------------------------------------------------------------ CUDA arm
def run_ours(args):
    import faulthandler
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # a stuck run prints where every thread is and exits, so that a driver is not left waiting; the stage log says how far
    # the rank came (stderr with IRONB_BENCH_VERBOSE=1, and gpurun_out/bench_rank<r>.log when that directory exists)
    faulthandler.enable()          # a host-side crash (SIGSEGV ...) prints the Python stacks to stderr
    faulthandler.dump_traceback_later(float(os.environ.get("IRONB_BENCH_WATCHDOG_S", "180")), exit=True)
    _t_start = time.perf_counter()
    _logf = None
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        _logf = open(os.path.join(ROOT, "gpurun_out", f"bench_rank{rank}.log"), "a", buffering=1)

    def stage(msg):
        STAGE[0] = msg
        line_ = f"[bench rank {rank} +{time.perf_counter() - _t_start:7.2f}s] {msg}\n"
        if _logf is not None:
            _logf.write(line_)
        if os.environ.get("IRONB_BENCH_VERBOSE"):
            sys.stderr.write(line_)

    stage("start")
    import torch
    import torch.distributed as dist
    stage("torch imported")
    # NVML is initialised before anything else (its start-up is slow and must not run next to the timed loop)
    sampler = ClockSampler(local) if (rank == 0 and not args.no_clocks) else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stage("process group up" if world > 1 else "device set")

    import iron_b200 as ib
    from iron_b200 import _lib
    from iron_b200.parallel import allreduce_flat, allreduce_gradients
    from oracle import iron_oracle as O   # fixture constants only (camera K / W2C); nothing is computed with it here
    lib = _lib.load()
    if args.tracer != "default":
        lib.ironb_set_trace_mode({"batched": 2, "fused": 0}[args.tracer])
    if args.gemm != "default":
        lib.ironb_set_gemm_mode({"tcgen05": 2, "tf32": 1, "ffma": 0}[args.gemm])
    _prev = lib.ironb_set_trace_mode(2)
    lib.ironb_set_trace_mode(_prev)
    tracer_impl = {2: "batched tcgen05 (fp16x2 split)", 0: "fused persistent fp32 FFMA"}[_prev]

    H, S = args.hidden, args.patch
    torch.manual_seed(0)
    nets = ib.init_rendering_network_dict("ggx")
    torch.manual_seed(0)
    sdf = ib.SDFNetwork(d_in=3, d_out=257, d_hidden=H, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                        geometric_init=True, weight_norm=True).to(dev)
    nets["point_light_network"].set_light(32.0)
    nets = {k: v.to(dev) for k, v in nets.items()}
    rend = ib.GGXColocatedRenderer(use_cuda=True)
    render_fn = ib.make_render_fn(rend)
    tracer = ib.RayTracer()
    tracer.collect_stats = True
    # weak scaling over patches: every rank traces and shades ITS OWN rays with its own target and eikonal samples
    # (shard_mode: the centre crop of its own view, or its own crop window of one view); rank 0 is always the canonical
    # configs[1] input.  The per-rank step and compute times are reported next to the max that defines `value`.
    crop_rank = int(os.environ.get("IRONB_BENCH_CROP_RANK", rank))      # diagnostic: run another rank's input on one GPU
    K_h = torch.tensor(O.FIXTURE_K, dtype=torch.float64).reshape(4, 4).float().pin_memory()
    W2C_h = torch.from_numpy(view_w2c(crop_rank)).float().pin_memory()
    ul = crop_corner(crop_rank, S) if not os.environ.get("IRONB_BENCH_SAME_CROP") else crop_corner(0, S)
    target_h = (torch.rand(S, S, 3, generator=torch.Generator().manual_seed(11 + rank)) * 0.5).pin_memory()
    eik_h = torch.empty(S * S // 2, 3).uniform_(-1.0, 1.0, generator=torch.Generator().manual_seed(12 + rank)).pin_memory()

    def make_cam(Kd, Wd):
        cam512 = ib.Camera(512, 512, Kd, Wd)
        return cam512.crop_region(S, S, ul_corner=ul)[0]

    cam = make_cam(K_h.to(dev), W2C_h.to(dev))
    target, eik = target_h.to(dev), eik_h.to(dev)
    params = [p for p in sdf.parameters()]
    for nm in ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network", "point_light_network"):
        params += list(nets[nm].parameters())
    n_params = sum(p.numel() for p in params)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    trace_ms = []

    use_graph = args.exec_mode == "graph" and args.shading == "dense" and not args.driver_defaults
    gs = None
    if use_graph:
        tracer.collect_stats = os.environ.get("IRONB_BENCH_STATS", "1") != "0"
        gs = ib.GraphedStage2Step(sdf, nets, tracer, render_fn, K_h, W2C_h, (S, S), S * S // 2, crop_ul=ul,
                                  time_tracer=os.environ.get("IRONB_BENCH_TIME_TRACER", "1") != "0", image_loss=args.loss,
                                  flat_grads=world > 1, grad_scale=1.0 / world)
        stage("graph captured")
        gs.step(target=target_h, eik_points=eik_h)
        torch.cuda.synchronize()
        stage("first replay done")

    mid_evs = []

    def step(cam_, target_, eik_, time_trace=False):
        if gs is not None:           # inputs are already in the graph's static buffers (cam_/target_/eik_ are those values)
            gs.graph.replay()
            if time_trace and world > 1:      # end of this rank's own work: what follows is the exchange (+ waiting for peers)
                mid = torch.cuda.Event(enable_timing=True)
                mid.record()
                mid_evs.append(mid)
            if world > 1:          # the graph's tail packed the gradients (x 1/world) into one buffer: one in-place all-reduce
                allreduce_flat(gs.flat_grad, world)
            return gs.loss, gs.results
        for p in params:
            p.grad = None
        if time_trace:   # time the dominant kernel (the tracer's four phase launches) on the launching stream
            orig = tracer.forward
            evs = []

            def timed(*a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = orig(*a, **k)
                e1.record()
                evs.append((e0, e1))
                return r
            tracer.forward = timed
        loss, res = ib.stage2_step(sdf, nets, tracer, render_fn, cam_, target_, eik_, fill_holes=args.driver_defaults,
                                   handle_edges=args.driver_defaults, dense_shading=(args.shading == "dense"),
                                   image_loss=args.loss)
        if time_trace:
            tracer.forward = orig
            trace_ms.append(evs)
        if world > 1:
            allreduce_gradients(params, world)      # the only collective of the path: flat fp32 gradient buffer, NCCL
        return loss, res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the cyclic garbage collector is paused over warm-up + timed loops: a generation-2 pass over the module graph landed
    # deterministically inside the 4th timed step (10-58 ms of host stall with the GPU idle)
    import gc
    gc.collect()
    gc.disable()
    # ---- warm-up: W steps (>= 3), extended to >= 0.5 s of continuous load.  ~60-80 ms after the GPU goes from idle to
    # busy the driver stalls launches once (1-2 ms alone, 40-120 ms when an NVML query is in flight at that moment --
    # measured, see profiles/README.md); that transition belongs to the warm-up, not to the timed steps.
    n_warm = 0
    t_w = time.perf_counter()
    min_warm_s = float(os.environ.get("IRONB_BENCH_MIN_WARMUP_S", "0.5"))    # 0 for ncu launch lists
    # Every step contains a collective (the gradient all-reduce), so ALL ranks must run the same number of warm-up steps:
    # the "is 0.5 s over?" decision is taken in chunks of 4 steps and agreed on by an all-reduce(MAX).  (Round 1 let every
    # rank consult its own clock: ranks left the loop after different step counts, the stragglers' all-reduce met the
    # others' barrier, and the 8-GPU run dead-locked in 2 of 3 launches -- SCALE_r01 N=8.)
    more = torch.zeros(1, dtype=torch.int32, device=dev)
    while True:
        for _ in range(4):
            step(cam, target, eik)
            n_warm += 1
        torch.cuda.synchronize()
        again = n_warm < max(args.warmup, 3) or time.perf_counter() - t_w < min_warm_s
        if world > 1:
            more.fill_(1 if again else 0)
            dist.all_reduce(more, op=dist.ReduceOp.MAX)
            again = bool(more.item())
        if not again:
            break
    barrier()
    stage(f"warm-up done ({n_warm} steps)")

    # ---- timed: K steps, per-step CUDA events, L2 flushed between steps
    if gs is None:
        tracer.last_stats = None
    l0 = lib.ironb_launch_count()
    evs = []
    hits = 0
    barrier()
    wall0 = time.perf_counter()
    epoch0 = time.time()
    sample_after = {args.steps // 3, (2 * args.steps) // 3, args.steps - 1}
    for i in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss, res = step(cam, target, eik, time_trace=True)
        e1.record()
        evs.append((e0, e1))
        if i in sample_after:
            # clocks under load: queried right after step i was enqueued (the GPU is still running it), in the gap between
            # two event pairs; with several ranks everybody re-aligns afterwards so a slow query cannot leak into a step
            if sampler:
                sampler.sample()
            if world > 1:
                barrier()
    barrier()
    stage("timed loop done")
    wall = time.perf_counter() - wall0
    epoch1 = time.time()
    launches = lib.ironb_launch_count() - l0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    my_ms = sum(step_ms)
    if gs is not None:
        # one graph replay per step: the library's kernels are graph nodes (counted at capture); the tracer's device time
        # comes from two external-event nodes inside the graph, readable for the last replay
        launches = gs.kernels_per_replay * args.steps
        tr_ms = None            # per-replay tracer times are read in the e2e loop below (each of its steps ends in a sync)
        stats = [v * args.steps for v in tracer.last_stats.cpu().tolist()]
        stats[5] = stats[5] // args.steps
    else:
        tr_ms = [sum(a.elapsed_time(b) for a, b in ev) for ev in trace_ms]
        stats = tracer.last_stats.cpu().tolist()
    hits = int(res["convergent_mask"].sum().item())
    t = torch.tensor([my_ms], dtype=torch.float64, device=dev)
    rank_ms = [my_ms / args.steps]
    rank_compute_ms = None
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rank_ms = [float(x.item()) / args.steps for x in allt]
        if len(mid_evs) == len(evs):      # replay only (own crop, before the all-reduce): shows the load imbalance between crops
            tc_ = torch.tensor([sum(a.elapsed_time(m) for (a, _), m in zip(evs, mid_evs))], dtype=torch.float64, device=dev)
            allc = [torch.zeros_like(tc_) for _ in range(world)]
            dist.all_gather(allc, tc_)
            rank_compute_ms = [round(float(x.item()) / args.steps, 4) for x in allc]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    rays_total = S * S * world * args.steps
    value = rays_total / (total_ms * 1e-3)

    # ---- e2e: same step through the public API from pinned host buffers (H2D + D2H inside the timed region)
    def e2e_step():
        if gs is not None:     # H2D into the graph's static buffers (patch, samples, camera matrices), replay, all-reduce
            gs.step(target=target_h, eik_points=eik_h, K=K_h, W2C=W2C_h)
            if world > 1:
                allreduce_flat(gs.flat_grad, world)
            return gs.loss
        # the camera is built from the HOST matrices (inverted on the host; K, W2C and their inverses are uploaded from
        # pinned memory), the target patch and the eikonal samples are copied from pinned host buffers
        tg, ek = target_h.to(dev, non_blocking=True), eik_h.to(dev, non_blocking=True)
        return step(make_cam(K_h, W2C_h), tg, ek)[0]

    float(e2e_step().item())   # untimed
    barrier()
    e2e_steps = args.steps
    t0 = time.perf_counter()
    tr_e2e = []
    for _ in range(e2e_steps):
        loss_host = float(e2e_step().item())          # D2H of the step's result
        if gs is not None and gs._tracer_events:      # the external event pair inside the graph: THIS replay's tracer time
            tr_e2e.append(gs.tracer_ms())
    if tr_ms is None:
        tr_ms = tr_e2e if tr_e2e else [0.0] * args.steps
    barrier()
    e2e_t = time.perf_counter() - t0
    te = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = S * S * world * e2e_steps / float(te.item())
    h2d = 2 * 4 * 16 * 4 + target_h.numel() * 4 + eik_h.numel() * 4   # two cameras (full view + crop) x 4 matrices, patch, samples
    d2h = 4 + (0 if args.shading == "dense" else 4)   # loss (+ the hit count the shading chunk reads back when compacting)

    stage("e2e loop done")
    gc.enable()
    clocks = sampler.stop(epoch0, epoch1) if sampler else None
    if rank == 0:
        pk = peaks()
        evals = stats[0] + stats[1] + stats[2]                       # SDF evaluations the tracer executed, all K steps
        flop_per_eval = 2.0 * (7 * H * H + H)                       # sdf-only evaluation
        tr_total_ms = sum(tr_ms)
        achieved = evals * flop_per_eval / (tr_total_ms * 1e-3) / 1e12 if tr_total_ms > 0 else 0.0
        peak = pk["bf16_tflops_sustained"]
        if _prev == 2:
            roof_kernel = ("mlp_h16_kernel (batched tcgen05 tracer: all 8 hidden SDF-MLP layers + sdf row per 128-row tile, "
                           "cluster of H/128 CTAs, two tiles in flight)")
            roof_mode = ("fp16x2 split on tcgen05 (3 kind::f16 MMAs per product, fp32-grade accuracy): the ceiling of this "
                         "arithmetic is peak/3")
            split_cost = 3.0
        else:   # --tracer fused: the persistent fp32-FFMA tracer; same FLOP count, reported against the same peak
            roof_kernel = "trace_kernel (fused persistent fp32-FFMA tracer, exact fp32 association)"
            roof_mode = "fp32 FFMA on the CUDA cores (diagnostic mode): the tensor peak is not attainable by this arithmetic"
            split_cost = 1.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "warmup_steps_run": n_warm,      # W steps, extended to >= 0.5 s of continuous load (see the warm-up loop)
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, S),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(te.item()) * 1e3 / e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"kernel": roof_kernel, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of ONE mlp_h16 launch of this workload (4,096 rows,
                         # H = 512) from the committed ncu --set full capture (profiles/r1l_mlp_h16_ncu.md, launch 0,
                         # cold cache: the fp16 weight copies, 7.3 MB, and the encoded points are read once from HBM; the
                         # activations live in L2).  The kernel's operand stream is L2 -> SM: 0.48 GB per launch.
                         # dram__bytes_read.sum + dram__bytes_write.sum of ONE mlp_h16 launch of this workload, NOT measured
                         # by this run: the constant is read from the committed ncu --set full capture named beside it
                         "traffic": (MLP_H16_TRAFFIC.get((H, S)) if _prev == 2 else None),
                         "traffic_source": MLP_H16_TRAFFIC_SOURCE,
                         "mode": roof_mode + "; achieved = ALGORITHMIC fp32 FLOPs of the evaluations executed / tracer time "
                                 "(whole tracer call incl. its state-machine kernels); peak = bf16 dense tensor, sustained, of "
                                 + pk["source"],
                         "frac_of_split_ceiling": achieved / (peak / split_cost) if peak else None,
                         "flop_per_eval": flop_per_eval, "evals_per_step": evals / args.steps,
                         "kernel_ms_per_step": tr_total_ms / max(len(tr_ms), 1),
                         "kernel_share_of_step": (tr_total_ms / len(tr_ms)) / (my_ms / args.steps) if (my_ms and tr_ms) else None},
            "tracer": {"evals_sphere": stats[0] / args.steps, "evals_sampler": stats[1] / args.steps,
                       "evals_bisect": stats[2] / args.steps, "sampler_rays": stats[3] / args.steps,
                       "root_rays": stats[4] / args.steps, "k_max": stats[5], "hits": hits, "rays": S * S,
                       "implementation": tracer_impl},
            "wall_ms_per_step_incl_flush": wall * 1e3 / args.steps, "grad_params": n_params,
            "rank_ms_per_step": [round(x, 4) for x in rank_ms], "rank_compute_ms_per_step": rank_compute_ms, "rank_crop_ul": [list(crop_corner(r, S)) for r in range(world)],
            "rank_view_deg": [45.0 * (r % 8) if shard_mode() == "views" else 0.0 for r in range(world)],
            "loss": loss_host,
            "step_ms": [round(x, 3) for x in step_ms], "tracer_ms": [round(x, 3) for x in tr_ms],
            "tracer_ms_source": ("one sample per replay of the e2e loop (external CUDA-event pair inside the graph)"
                                 if gs is not None else "CUDA-event pair around every tracer call of the timed steps"),
        }
        if world == 1 and not args.no_cpu:
            line["ggx_roofline"] = ggx_microbench(dev, pk)
            # the reference path as eager PyTorch on this GPU (the competitor on equal hardware), then on the host cores
            line["cuda_eager_baseline"] = cuda_eager_baseline(H, S, dev)
            line["cpu_baseline"] = cpu_baseline(H, S)
            ce, cb = line["cuda_eager_baseline"], line["cpu_baseline"]
            ce["speedup_of_this_library"] = {"device_timed": value / ce["value"], "e2e": e2e_value / ce["value"]}
            rel = lambda a, b: abs(a - b) / max(abs(b), 1e-30)
            line["loss_check"] = {      # same weights, crop, target and eikonal samples in all three
                "loss_gpu": loss_host, "loss_reference_cuda_eager": ce["loss"], "rel_err_vs_cuda_eager": rel(loss_host, ce["loss"]),
                "loss_oracle_cpu": cb["loss"] if cb["rays"] == S * S else None,
                "rel_err_vs_oracle_cpu": rel(loss_host, cb["loss"]) if cb["rays"] == S * S else None}
        emit(line)
    faulthandler.cancel_dump_traceback_later()
    # orderly teardown while the CUDA context is alive: the captured graph (its memory pool, streams and external events)
    # first, then the process group; the interpreter then exits normally with the process's real status
    if gs is not None:
        gs.close()
    del gs
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    stage("done")
    sys.stdout.flush()
    sys.stderr.flush()


def run_render(args):
    """--workload render1024 (BASELINE configs[2]): render_surface.py's full-frame forward render -- the 512x512 fixture view
    resized to 1024 x 1024 (1,048,576 rays), render_camera(is_training=False): sphere tracing + dense sampling + bisection in
    <= 50,000-ray tracer calls, get_all + materials + GGX on the compacted hits in <= 320,000-point chunks, no backward.
    Single GPU.  `value` = rays/s device-timed; `e2e` adds the camera upload and the D2H copy of the rendered image."""
    import torch
    import iron_b200 as ib
    from iron_b200 import _lib
    from oracle import iron_oracle as O
    lib = _lib.load()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    H = args.hidden
    torch.manual_seed(0)
    nets = ib.init_rendering_network_dict("ggx")
    torch.manual_seed(0)
    sdf = ib.SDFNetwork(d_in=3, d_out=257, d_hidden=H, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                        geometric_init=True, weight_norm=True).to(dev)
    nets["point_light_network"].set_light(32.0)
    nets = {k: v.to(dev) for k, v in nets.items()}
    render_fn = ib.make_render_fn(ib.GGXColocatedRenderer(use_cuda=True))
    tracer = ib.RayTracer()
    tracer.collect_stats = True
    K_h = torch.tensor(O.FIXTURE_K, dtype=torch.float64).reshape(4, 4).float().pin_memory()
    W2C_h = torch.tensor(O.FIXTURE_W2C, dtype=torch.float64).reshape(4, 4).float().pin_memory()
    R = 1024

    def make_cam():
        return ib.Camera(512, 512, K_h.to(dev, non_blocking=True), W2C_h.to(dev, non_blocking=True)).resize(R / 512.0)[0]

    cam = make_cam()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pin = torch.empty(R, R, 3, dtype=torch.float32).pin_memory()

    def frame(c):
        with torch.no_grad():
            return ib.render_camera(c, sdf, tracer, nets, render_fn, fill_holes=False, handle_edges=False, is_training=False)

    for _ in range(max(args.warmup, 3)):
        res = frame(cam)
    torch.cuda.synchronize()
    l0 = lib.ironb_launch_count()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = frame(cam)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    launches = lib.ironb_launch_count() - l0
    ms = [a.elapsed_time(b) for a, b in evs]
    t = sum(ms) / len(ms)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = frame(make_cam())
        pin.copy_(res["color"], non_blocking=True)
        torch.cuda.synchronize()
    e2e_t = (time.perf_counter() - t0) / args.steps
    hits = int(res["convergent_mask"].sum().item())
    line = {"metric": "traced+shaded rays/sec forward (1024x1024 full-frame render, 8x512 SDF MLP, GGX)", "value": R * R / (t * 1e-3),
            "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[2]: render_surface.py full-frame forward render 1024x1024 (the 512x512 fixture view "
                                   "resized x2): sphere tracing + dense sampling + bisection + get_all + materials + GGX, no backward",
                       "sdf_mlp": f"8x{H}, PE L=6, skip@4, softplus(100), weight-norm", "rays": R * R, "hits": hits,
                       "chunking": "<= 50,000 rays per tracer call, <= 320,000 hit points per shading chunk (the reference's)",
                       "l2": "256 MiB flush between frames", "execution": "eager (the hit count of every shading chunk is read back)"},
            "e2e": {"value": R * R / e2e_t, "unit": UNIT, "h2d_bytes_per_step": 2 * 64, "d2h_bytes_per_step": R * R * 3 * 4,
                    "ms_per_step": e2e_t * 1e3},
            "gpu_launches": int(launches), "step_ms": [round(x, 2) for x in ms]}
    if not args.no_cpu:
        # the reference path as eager PyTorch on this GPU, same frame (one warm-up + one timed frame: ~10 s each)
        import torch.backends.cuda
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        st = oracle_setup(H, 64, 0, device=dev)
        Oc, sdf_p, mats, light = st[0], st[1], st[2], st[3]
        fx = Oc.OCamera.fixture()
        ocam = Oc.OCamera(512, 512, fx.K.to(dev), fx.W2C.to(dev)).resize(R / 512.0)
        with torch.device(dev), torch.no_grad():
            Oc.render_camera(sdf_p, mats, light.detach(), ocam.crop(256, 256, (384, 384)), handle_edges=False)   # warm-up
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ores = Oc.render_camera(sdf_p, mats, light.detach(), ocam, handle_edges=False)
            e1.record()
            torch.cuda.synchronize()
        tm = e0.elapsed_time(e1)
        m, mr = res["convergent_mask"], ores["convergent_mask"]
        both = m & mr
        line["cuda_eager_baseline"] = {
            "value": R * R / (tm * 1e-3), "unit": UNIT, "ms_per_step": tm, "steps": 1,
            "what": "the reference path (oracle restatement) as eager PyTorch on cuda:0, fp32, allow_tf32=False, same frame",
            "speedup_of_this_library": line["value"] / (R * R / (tm * 1e-3)),
            "mask_agreement": float((m == mr).float().mean()),
            "rgb_within_1e-3": float(((res["color"] - ores["color"]).abs().amax(-1)[both] <= 1e-3).float().mean())}
    emit(line)


def run_neus(args):
    """--workload neus: one stage-1 training step of render_volume.py:230-290 -- NeuSRenderer.render on 512 rays
    (confs/*_iron.conf: 64 + 64 samples in 4 up-sample steps, 32 outside samples through the background NeRF, perturb on,
    SDF MLP 8 x 256, colour MLP 8 x 256 with its skip layer), the stage-1 loss (L1 colour + 0.1 eikonal + 0.1 BCE mask) and
    its backward.  Single GPU, eager (the hierarchical sampler is a sequence of small tensor ops).  `value` = rays/s
    device-timed; `e2e` adds the H2D copy of the ray batch from pinned memory and the D2H read of the loss."""
    import torch
    import iron_b200 as ib
    from iron_b200 import _lib
    from oracle import iron_oracle as O
    lib = _lib.load()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    H, B = 256, 512
    torch.manual_seed(0)
    sdf = ib.SDFNetwork(d_in=3, d_out=257, d_hidden=H, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                        geometric_init=True, weight_norm=True)
    color = ib.RenderingNetwork(d_feature=256, mode="idr", d_in=9, d_out=3, d_hidden=H, n_layers=8, skip_in=[4],
                                weight_norm=True, multires=10, multires_view=4, squeeze_out=True)
    devn = ib.SingleVarianceNetwork(0.3)
    nerf = ib.NeRF(D=8, W=256, d_in=4, d_in_view=3, multires=10, multires_view=4, output_ch=4, skips=[4], use_viewdirs=True)
    p = lambda mod: {k: v.detach().clone().to(dev).requires_grad_(True) for k, v in mod.state_dict().items()}
    sdf_p, color_p, nerf_p = p(sdf), p(color), p(nerf)
    var = devn.variance.detach().clone().to(dev).requires_grad_(True)
    for m in (sdf, color, devn, nerf):
        m.to(dev)
    gen = torch.Generator().manual_seed(9)
    o = torch.randn(B, 3, generator=gen)
    o = o / o.norm(dim=-1, keepdim=True) * 2.0
    d = (torch.rand(B, 3, generator=gen) - 0.5) * 1.2 - o
    d = d / d.norm(dim=-1, keepdim=True)
    mid = -(o * d).sum(-1, keepdim=True)
    host = [t.contiguous().pin_memory() for t in (o, d, mid - 1.0, mid + 1.0, torch.rand(B, 3, generator=gen),
                                                  (torch.rand(B, 1, generator=gen) > 0.4).float())]
    t_rand, t_out = torch.rand(B, 1, generator=gen).to(dev), torch.rand(B, 32, generator=gen).to(dev)
    bg = torch.ones(1, 3, device=dev)
    ren = ib.NeuSRenderer(nerf, sdf, devn, color, n_samples=64, n_importance=64, n_outside=32, up_sample_steps=4, perturb=1.0)
    params = [q for m in (sdf, color, devn, nerf) for q in m.parameters()]

    def loss_of(out, target, mask):
        color_loss = (out["color_fine"] - target).abs().sum() / target.shape[0]
        mask_loss = torch.nn.functional.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask)
        return color_loss + 0.1 * out["gradient_error"] + 0.1 * mask_loss

    def step(batch):
        for q in params:
            q.grad = None
        draws = [t_rand, t_out]
        ren.rand_fn = lambda shape: draws.pop(0).reshape(shape)           # the same uniform numbers in both arms
        out = ren.render(batch[0], batch[1], batch[2], batch[3], background_rgb=bg, cos_anneal_ratio=0.5)
        loss = loss_of(out, batch[4], batch[5])
        loss.backward()
        return loss.detach()

    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False          # the background NeRF's cuBLAS GEMMs stay fp32
    torch.backends.cudnn.allow_tf32 = False
    on_dev = [t.to(dev) for t in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gs = None
    if args.exec_mode == "graph":
        # built BEFORE any eager backward through these parameters (their AccumulateGrad nodes remember the stream they were
        # created on; a node created on the legacy stream cannot be captured)
        gs = ib.GraphedNeusStep(ren, B, loss_of, background_rgb=bg, cos_anneal_ratio=0.5)
    for _ in range(max(args.warmup, 3)):
        step(on_dev)
    torch.cuda.synchronize()
    l0 = lib.ironb_launch_count()
    ms = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = step(on_dev)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    launches = lib.ironb_launch_count() - l0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss_host = float(step([t.to(dev, non_blocking=True) for t in host]).item())
    e2e_t = (time.perf_counter() - t0) / args.steps
    loss_eager_arm = loss_host
    sdf_grads_eager_arm = {k: q.grad.clone() for k, q in sdf.named_parameters()}
    t_med = statistics.median(ms)
    eager_ms, eager_e2e = list(ms), e2e_t
    graph_info = None
    if gs is not None:
        # the same step as ONE CUDA-graph replay (iron_b200.GraphedNeusStep); the perturbation draws then come from torch's
        # graph-safe generator (fresh numbers per replay), so this arm's loss differs from the eager arm's by the draws
        for _ in range(max(args.warmup, 3)):
            gs.step(*on_dev)
        torch.cuda.synchronize()
        ms = []
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gs.graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            loss_host = float(gs.step(*host).item())
        e2e_t = (time.perf_counter() - t0) / args.steps
        t_med = statistics.median(ms)
        launches = gs.kernels_per_replay * args.steps
        graph_info = {"eager_ms_per_step": statistics.median(eager_ms), "eager_e2e_ms_per_step": eager_e2e * 1e3,
                      "kernels_per_replay": gs.kernels_per_replay}
        gs.close()
    line = {"metric": "stage-1 NeuS volume-rendered rays/sec fwd+bwd (512 rays x (128 + 32) sections, SDF MLP 8x256)",
            "value": B / (sum(ms) / len(ms) * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": sum(ms) / len(ms), "ms_per_step_median": t_med, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SURVEY 8 f-4: stage-1 NeuS training step (render_volume.py:230-290), 512 rays, n_samples 64 + "
                                   "n_importance 64 (4 up-sample steps), n_outside 32 (background NeRF 8x256), perturb on, white "
                                   "background, cos_anneal 0.5; loss = L1 colour + 0.1 eikonal + 0.1 BCE mask; backward to all four "
                                   "networks", "sdf_mlp": "8x256, PE L=6, skip@4, softplus(100), weight-norm",
                       "colour_mlp": "8x256, PE L=10 / view L=4, skip@4, weight-norm", "rays": B, "sdf_points_per_step": B * (64 + 48 + 128),
                       "l2": "256 MiB flush between steps",
                       "execution": "one CUDA-graph replay per step (iron_b200.GraphedNeusStep)" if graph_info else "eager"},
            "e2e": {"value": B / e2e_t, "unit": UNIT, "h2d_bytes_per_step": sum(t.numel() * 4 for t in host), "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_t * 1e3},
            "gpu_launches": int(launches), "loss": loss_host, "step_ms": [round(x, 3) for x in ms], "graph": graph_info}
    if not args.no_cpu:
        # the reference's renderer (oracle restatement, pinned to the real NeuSRenderer by tests/golden/neus.npz) as eager
        # PyTorch on this GPU: same weights, rays and uniform numbers
        def ref_step():
            for q in list(sdf_p.values()) + list(color_p.values()) + list(nerf_p.values()) + [var]:
                q.grad = None
            with torch.device(dev):
                out = O.neus_render(sdf_p, color_p, var, nerf_p, on_dev[0], on_dev[1], on_dev[2], on_dev[3], n_samples=64,
                                    n_importance=64, n_outside=32, up_sample_steps=4, t_rand=t_rand, t_rand_outside=t_out,
                                    background_rgb=bg, cos_anneal_ratio=0.5)
                l = loss_of(out, on_dev[4], on_dev[5])
                l.backward()
            return l.detach()
        for _ in range(2):
            ref_step()
        torch.cuda.synchronize()
        rms = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rl = ref_step()
            e1.record()
            torch.cuda.synchronize()
            rms.append(e0.elapsed_time(e1))
        tr = statistics.median(rms)
        gmax = max(float((sdf_grads_eager_arm[k] - sdf_p[k].grad).norm() / sdf_p[k].grad.norm().clamp_min(1e-30)) for k, _ in sdf.named_parameters())
        line["cuda_eager_baseline"] = {
            "value": B / (tr * 1e-3), "unit": UNIT, "ms_per_step": tr, "steps": 5, "loss": float(rl),
            "what": "the reference's NeuSRenderer (oracle restatement) + autograd as eager PyTorch on cuda:0, fp32, allow_tf32=False, "
                    "same weights / rays / uniform numbers",
            "speedup_of_this_library": (B / (t_med * 1e-3)) / (B / (tr * 1e-3)),
            "loss_rel_err": abs(loss_eager_arm - float(rl)) / abs(float(rl)), "worst_sdf_gradient_rel_l2": gmax,
            "parity_note": "loss / gradients compared on the eager arm, which uses the same uniform numbers as the baseline"}
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    emit(line)


_REAL_STDOUT = None
STAGE = ["not started"]      # the last stage run_ours reached (named in the failure message)


def emit(line: dict):
    """The ONE JSON line goes to the process's real stdout; everything else written to fd 1 meanwhile (NCCL prints
    'NCCL version ...' there) has been routed to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)             # C-level writers to stdout (NCCL banner) must not precede the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--hidden", type=int, default=512)
    ap.add_argument("--patch", type=int, default=64)
    ap.add_argument("--loss", default="reference", choices=["reference", "l2"],
                    help="reference: PyramidL2 + SSIM + roughness range (render_surface.py:594-613), in every arm; l2: round 1's plain L2")
    ap.add_argument("--workload", default="step", choices=["step", "render1024", "neus"],
                    help="step: the stage-2 training step (configs[1]; --patch 256 = configs[3]); render1024: configs[2], the "
                         "full-frame 1024x1024 forward render; neus: the stage-1 NeuS training step (SURVEY 8 f-4)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--driver-defaults", action="store_true",
                    help="also run hole filling + edge sampling (the reference drivers' fill_holes=True, handle_edges=True)")
    ap.add_argument("--exec", dest="exec_mode", default="graph", choices=["graph", "eager"],
                    help="graph: the whole step is one CUDA-graph replay (iron_b200.GraphedStage2Step; needs --shading dense); "
                         "eager: ~700 launches per step from Python")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample clocks (diagnosing sampler interference)")
    ap.add_argument("--shading", default="dense", choices=["dense", "compact"],
                    help="dense: shade every ray and mask (no hit-count read-back, host runs ahead of the tracer); compact: the "
                         "reference's order (compact the hits first: one host sync per step)")
    ap.add_argument("--tracer", default="default", choices=["default", "batched", "fused"])
    ap.add_argument("--gemm", default="default", choices=["default", "tcgen05", "tf32", "ffma"])
    args = ap.parse_args()
    IMAGE_LOSS[0] = args.loss
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "render1024":
        run_render(args)
        return
    if args.workload == "neus":
        run_neus(args)
        return
    guarded_run_ours(args)


def guarded_run_ours(args):
    """run_ours with its failure made visible: the traceback goes to stderr (and to gpurun_out/bench_rank<r>.err when that
    directory exists), then the process leaves at once with a non-zero code.  A rank that raised must not enter the
    interpreter's teardown: destroying a NCCL process group whose peers are still inside a collective blocks, the rank
    lingers with an idle GPU until the watchdog fires, and the launcher's summary then hides which rank failed first and
    why (SCALE_r01 N=8).  With an immediate exit torchrun sees the first failure, names it, and stops the peers."""
    import traceback
    rank = os.environ.get("RANK", "0")
    try:
        run_ours(args)
    except BaseException as e:      # noqa: BLE001 -- includes KeyboardInterrupt / SystemExit from the watchdog
        if isinstance(e, SystemExit) and e.code in (0, None):
            raise
        msg = f"bench.py: rank {rank} FAILED after stage '{STAGE[0]}'\n{traceback.format_exc()}"
        sys.stderr.write(msg)
        sys.stderr.flush()
        try:
            d = os.path.join(ROOT, "gpurun_out")
            if os.path.isdir(d):
                with open(os.path.join(d, f"bench_rank{rank}.err"), "a") as f:
                    f.write(msg)
            ef = os.environ.get("TORCHELASTIC_ERROR_FILE")      # torchrun prints this under "Root Cause"
            if ef:
                with open(ef, "w") as f:
                    json.dump({"message": {"message": f"{type(e).__name__}: {e}", "extraInfo":
                                           {"py_callstack": traceback.format_exc(), "timestamp": str(int(time.time()))}}}, f)
        except Exception:
            pass
        os._exit(1)


if __name__ == "__main__":
    main()
