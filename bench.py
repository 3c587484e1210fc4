#!/usr/bin/env python
"""bench.py -- views/sec forward+backward of the GOI rasterizer hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2]

One "step" = one full training-view pass per GPU: render() forward (preprocess, binning, depth sort,
composite of RGB + S semantic channels + depth + alpha) + backward of a fixed linear pseudo-loss
L = sum(w * outputs) down to every per-Gaussian input (means, SH, semantics, opacity, scales,
rotations).  Workload at N=1 = BASELINE.json configs[1] stand-in ("c2": 1M synthetic Gaussians,
1600x1000, 16 semantic channels; the pretrained 'garden' scene is not on the box).
N > 1: one process per GPU, each rank renders its own view per step against a full parameter replica
(weak scaling) and the per-Gaussian gradients are summed with one NCCL all-reduce per step.

JSON line (rank 0):
  value     views/s with all inputs resident in HBM (CUDA events, max over ranks)
  e2e       views/s through the same public API with the per-view inputs (camera + the pixel-space
            supervision = loss-weight images) coming from pinned HOST memory every step and the loss
            read back to the host; Gaussian parameters are model state and stay in HBM in both arms,
            as in the reference's train.py
  roofline  forward+backward composite kernels: algorithmic bytes B_comp / their CUDA-event
            time inside the timed region, against the measured HBM peak
  cpu_baseline  the CPU oracle port timed on the host cores on a bounded sample (reported only)
`--impl reference` times the reference's OWN CUDA kernels + host glue (oracle/_ref, built from
/root/reference by oracle/build.py) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goi-hyperplane_b200"))

import torch  # noqa: E402

CONFIGS = {
    # name: (P, W, H, S, seed)   SURVEY.md section 8d
    "c1": (10_000, 256, 256, 10, 0),
    "c2": (1_000_000, 1600, 1000, 16, 1),
    "c3": (1_000_000, 800, 600, 32, 2),
    "c4": (5_000_000, 1920, 1080, 16, 3),      # BASELINE config 4: 64 views over 8 GPUs = --gpus 8 --views-per-step 8
    "c5_4": (1_000_000, 1280, 720, 4, 4), "c5_8": (1_000_000, 1280, 720, 8, 4),
    "c5_16": (1_000_000, 1280, 720, 16, 4), "c5_32": (1_000_000, 1280, 720, 32, 4),
    "c5_64": (1_000_000, 1280, 720, 64, 4),
}
N_VIEWS = 8      # distinct camera poses cycled through (small yaw/pitch jitter around the scene axis)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt, self.proc = index, [], threading.Event(), None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                if self._stop_evt.is_set():
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc:
            self.proc.terminate()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        try:
            sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = int(self.samples[0][1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for i, n in enumerate(names):
                if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples):
                    out["reasons"].append(n)
        except Exception:
            pass
        return out


def make_views(W, H, device):
    from goi_b200.scenes import SyntheticCamera
    cams = []
    for v in range(N_VIEWS):
        yaw, pitch = math.radians(2.0 * math.sin(v * 0.9)), math.radians(1.5 * math.cos(v * 1.7))
        cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
        Ry = torch.tensor([[cy, 0, sy, 0], [0, 1, 0, 0], [-sy, 0, cy, 0], [0, 0, 0, 1.0]])
        Rx = torch.tensor([[1, 0, 0, 0], [0, cp, -sp, 0], [0, sp, cp, 0], [0, 0, 0, 1.0]])
        cams.append(SyntheticCamera(W, H, math.radians(60.0), Rx @ Ry, device=device))
    return cams


def bind_to_gpu_numa(index: int):
    """Pin this rank (and therefore the first-touch placement of its pinned staging buffers) to the CPUs that are local
    to GPU `index` (sysfs local_cpulist of the GPU's PCI function).  Returns a short description, or None."""
    try:
        pr = torch.cuda.get_device_properties(index)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(path + "/local_cpulist") as f:
            spec = f.read().strip()
        with open(path + "/numa_node") as f:
            node = int(f.read().strip())
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        pass
    return None


def build_workload(config, dev):
    """(gaussians, cameras, background) of a config: SURVEY.md section 8d.  c4 is the object-centric orbit scene (64
    cameras on a circle looking at the origin), the others the frustum-filling scene with 8 jittered poses."""
    from goi_b200.scenes import make_orbit_scene, make_scene
    P, W, H, S, seed = CONFIGS[config]
    if config == "c4":
        g, cams, bg = make_orbit_scene(P, W, H, S, 64, seed)
        return g.to(dev), [c.to(dev) for c in cams], bg.to(dev)
    g, _, bg = make_scene(P, W, H, S, seed)
    return g.to(dev), make_views(W, H, dev), bg.to(dev)


def ncu_reference(config):
    """DRAM bytes and executed warp instructions of the two composite kernels from the committed `ncu --set full`
    capture (profiles/traffic.json), for roofline.traffic and the issue-slot companion bound; None if the capture
    is of another config."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        if d.get("config") != config:
            return None
        ks = d["kernels"].values()
        return {"dram_bytes": sum(k["dram_bytes"] for k in ks),
                "warp_instructions": sum(k["warp_instructions"] for k in ks), "source": d["source"]}
    except Exception:
        return None


def metric_name(config):
    """BASELINE.json's metric string for the config it is quoted on (c2); the same sentence with the config's own
    sizes for the others, so a line can never be mistaken for the headline."""
    P, W, H, S, _ = CONFIGS[config]
    pm = f"{P // 1_000_000}M" if P % 1_000_000 == 0 else f"{P // 1000}k"
    return f"views/sec fwd+bwd @{pm} Gaussians,{W}x{H},{S}ch; HBM GB/s vs roofline"


def config_dict(args, world, vps):
    """`config` of the JSON line: identical (keys and values) for both arms."""
    P, W, H, S, _ = CONFIGS[args.config]
    return {"workload": f"{args.config}: P={P} Gaussians, {W}x{H}, S={S} semantic channels, SH degree 3, "
                        f"fwd+bwd of all four outputs, {64 if args.config == 'c4' else N_VIEWS} camera poses",
            "views_per_step": world * vps, "views_per_gpu_per_step": vps,
            "parallelism": f"view-dp{world}, one all-reduce of the per-Gaussian gradients per step",
            "l2_policy": "inputs larger than L2 (300 MB parameters + 98 MB sort buffers per view)"}


def run_dict(args, num_rendered_mean):
    """What differs between the arms by construction (kept OUT of `config`, which is identical for both): this build
    culls provably empty tile instances and does not read the instance count back per view."""
    return {"num_rendered_mean": round(num_rendered_mean),
            "host_sync": "per view (num_rendered read-back)" if (args.sync_binning or args.impl == "reference")
                         else "none per view (instance count stays on the device; status words checked per step)"}


def b_comp(R, W, H, S):
    """algorithmic bytes of forward + backward composite for one view."""
    T = ((W + 15) // 16) * ((H + 15) // 16)
    return 16 * T + R * (128 + 12 * S) + W * H * (8 * S + 52)


def run_ours(args):
    import torch.distributed as dist
    from diff_gaussian_rasterization import _C
    from gaussian_renderer import render
    from goi_b200.scenes import PipeFlags, make_loss_weights
    from goi_b200 import view_parallel as vp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = bind_to_gpu_numa(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _C.lib()        # hard error if the CUDA library is missing: no fallback path

    P, W, H, S, seed = CONFIGS[args.config]
    g, cams, bg = build_workload(args.config, dev)
    g = g.requires_grad_(True)
    n_cams = len(cams)
    pipe = PipeFlags()
    w_dev = make_loss_weights(S, W, H, seed, device=dev)
    # the backward writes the parameter gradients straight into slices of ONE flat buffer (all-reduce operand)
    arena = vp.GradArena({"means3D": g.get_xyz, "opacities": g.get_opacity, "scales": g.get_scaling,
                          "rotations": g.get_rotation, "sh": g.get_features, "semantics": g.get_semantics})

    outs = ("render", "semantics", "depth", "alpha")

    vps = args.views_per_step

    def step(i, weights, after_view=None):
        """One data-parallel step: `vps` training views per rank, each a render() forward + backward from
        dL/d(outputs) = the weights of the linear pseudo-loss L = sum(w * out) (its exact gradient) down to
        every per-Gaussian parameter; views after the first ADD their gradients in place to the flat buffer
        (goi_bwd_out.accumulate); one NCCL all-reduce of that buffer ends the step (BASELINE config 4 shards
        64 views over 8 GPUs the same way: 8 views per rank per all-reduce)."""
        for v in range(vps):
            view = i * vps + v                       # this rank's running view counter
            cam = cams[(view * world + rank) % n_cams]
            wv = weights(view) if callable(weights) else weights
            arena.clear_grads()
            out = render(cam, g, pipe, bg)
            with arena.accumulating(v > 0):
                torch.autograd.backward([out[k] for k in outs], [wv[k] for k in outs])
            if after_view is not None:
                after_view(view, out, wv)
        if world > 1:
            arena.all_reduce()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---------------- device-resident arm ----------------
    step(0, w_dev)                  # (seeds the instance-count estimate through the synchronous path)
    # training-loop mode: no num_rendered read-back per view (goi_forward_async); every view's status word is examined
    # -- overflow of the binning capacity would invalidate the run -- when the timed region ends
    _C.set_async_binning(not args.sync_binning, dev)
    for i in range(args.warmup):
        step(i, w_dev)
    _C.check_async(dev, wait=True)
    _C.timing_enable(True)
    stage_acc, rs = {}, []
    launches0 = _C.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    def resident_step(i):
        step(i, w_dev)
        _C.check_async(dev, wait=False)    # examines the status words that have already landed; never blocks
        rs.append(_C.num_rendered())       # (of the most recent examined view; the poses differ by a few degrees)

    ms_total = timed(resident_step, args.steps)
    _C.check_async(dev, wait=True)         # raises if any view of the timed region overflowed its binning capacity
    # per-stage CUDA-event times of the timed region's views (ring of the last 64), read after the region
    for k, v in _C.timing_read().items():
        stage_acc[k] = max(v, 0.0) * args.steps          # mean ms per VIEW x steps
    launches = _C.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else {}
    _C.timing_enable(False)
    value = world * vps * args.steps / (ms_total / 1e3)

    # ---------------- end-to-end arm: per-view inputs from pinned host memory ----------------
    # The per-view supervision travels the way a trainer holds it: the RGB map as uint8, the feature / depth / alpha maps
    # as fp16 (39 B per pixel instead of 84), and is expanded to the f32 loss-weight images on the device.
    def compact(wd):
        return {"render": ((wd["render"] + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8),
                "semantics": wd["semantics"].half(), "depth": wd["depth"].half(), "alpha": wd["alpha"].half()}
    host_w = [{k: v.cpu().pin_memory() for k, v in compact(make_loss_weights(S, W, H, seed + j)).items()} for j in range(2)]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host_w[0].values()) + (16 + 16 + 3 + 3) * 4
    copy_stream = torch.cuda.Stream(device=dev)
    # ring of NSLOT device slots, copies issued two views ahead: a view's copy then has two view periods to land
    NSLOT = 3
    slots = [{k: torch.empty_like(v, device=dev) for k, v in host_w[0].items()} for _ in range(NSLOT)]
    expanded = {k: torch.empty_like(v) for k, v in w_dev.items()}      # f32 images the backward reads (one set: the
    ready = [torch.cuda.Event() for _ in range(NSLOT)]                 # compute stream expands right before a view)
    freed = [torch.cuda.Event() for _ in range(NSLOT)]
    loss_host = [torch.zeros(2).pin_memory() for _ in range(NSLOT)]

    n_e2e_views = args.steps * vps

    _dbg = os.environ.get("GOI_BENCH_E2E_DEBUG", "")      # measurement experiments only: "nocopy", "noloss"

    def prefetch(view):
        s = view % NSLOT
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            if "nocopy" not in _dbg:
                for k in slots[s]:
                    slots[s][k].copy_(host_w[view % 2][k], non_blocking=True)
            ready[s].record(copy_stream)

    def view_inputs(view):
        """Called right before a view's forward: start the NEXT view's host->device copy (it overlaps this
        view's compute), then make the compute stream wait for this view's own copy."""
        v = view % n_e2e_views
        if v == 0:
            prefetch(view)
            if n_e2e_views > 1:
                prefetch(view + 1)
        if v + 2 < n_e2e_views:
            prefetch(view + 2)
        torch.cuda.current_stream().wait_event(ready[view % NSLOT])
        sl = slots[view % NSLOT]
        expanded["render"].copy_(sl["render"])
        expanded["render"].mul_(1.0 / 127.5).sub_(1.0)
        for k in ("semantics", "depth", "alpha"):
            expanded[k].copy_(sl[k])
        freed[view % NSLOT].record()              # the compact slot may now be overwritten by the copy stream
        return expanded

    def view_done(view, out, wv):
        # D2H read of the view's result: the pseudo-loss value and the gradient norm of the semantic field
        # (asynchronous, like a trainer's logging: ordered on the stream, drained at the end of the region)
        if "noloss" not in _dbg:
            loss = torch.dot(out["semantics"].detach().reshape(-1), wv["semantics"].reshape(-1))   # one pass, no temp
            loss_host[view % NSLOT].copy_(torch.stack([loss, torch.linalg.vector_norm(arena.slots['semantics'])]),
                                          non_blocking=True)

    def e2e_step(i):
        step(i, view_inputs, view_done)

    for ev in freed:
        ev.record()
    n_e2e_views = min(2, args.warmup) * vps
    for i in range(min(2, args.warmup)):
        e2e_step(i)
    torch.cuda.synchronize()
    for ev in freed:
        ev.record()
    n_e2e_views = args.steps * vps
    if "stages" in _dbg:
        _C.timing_enable(True)
    ms_e2e = timed(e2e_step, args.steps)
    _C.check_async(dev, wait=True)
    if "stages" in _dbg:
        print("e2e stages", {k: round(v, 4) for k, v in _C.timing_read().items()}, file=sys.stderr)
        _C.timing_enable(False)
    e2e_value = world * vps * args.steps / (ms_e2e / 1e3)

    if rank != 0:
        return
    peak, peak_src = peaks()
    R = sum(rs) / max(len(rs), 1)
    t_comp = (stage_acc.get("composite_fwd", 0.0) + stage_acc.get("composite_bwd", 0.0)) / args.steps / 1e3
    bytes_comp = b_comp(R, W, H, S)
    achieved = bytes_comp / t_comp / 1e9 if t_comp > 0 else 0.0
    line = {
        "metric": metric_name(args.config),
        "value": round(value, 3), "unit": "views/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, world, vps),
        "run": run_dict(args, R),
        "clocks": clocks,
        "ms_per_view": round(ms_total / args.steps / vps, 4),
        "e2e": {"value": round(e2e_value, 3), "unit": "views/s", "h2d_bytes_per_step": h2d_bytes * vps,
                "d2h_bytes_per_step": 8 * vps, "ms_per_step": round(ms_e2e / args.steps, 4),
                "supervision": "per view from pinned host memory: uint8 RGB + fp16 feature / depth / alpha loss-weight "
                               "maps (39 B per pixel at S=16), expanded to f32 on the device; camera matrices; the loss "
                               "and the semantic gradient norm are read back",
                "host_affinity": affinity},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                     "kernel": "k_composite_fwd + k_composite_bwd",
                     "algorithmic_bytes_per_view": int(bytes_comp), "kernel_ms_per_view": round(t_comp * 1e3, 4)},
        "stages_ms_per_view": {k: round(v / args.steps, 4) for k, v in stage_acc.items()},
    }
    ref = ncu_reference(args.config)
    if ref is not None and t_comp > 0:
        # measured DRAM traffic of the same two launches (ncu), and the bound that actually limits them: warp
        # instruction issue (4 schedulers x 1 instruction/clk per SM).  ncu's instruction count / this run's time.
        line["roofline"]["traffic"] = ref["dram_bytes"]
        line["roofline"]["traffic_source"] = ref["source"]
        sm_clk = (clocks.get("sm_mhz") or 1965) * 1e6
        issue_peak = 148 * 4 * sm_clk
        line["roofline"]["issue_bound"] = {
            "achieved": round(ref["warp_instructions"] / t_comp / 1e9, 1), "peak": round(issue_peak / 1e9, 1),
            "unit": "G warp-instructions/s", "frac": round(ref["warp_instructions"] / t_comp / issue_peak, 4),
            "note": "composites are issue-bound, not HBM-bound (DESIGN.md section 4): executed warp instructions "
                    "(ncu) over this run's composite time, against 148 SMs x 4 schedulers x the SM clock"}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line), flush=True)


def cpu_baseline(args):
    """The CPU oracle port on the host cores, on a bounded sample of the same workload: the same scene
    generator at 1/4 of the pixels and Gaussians (same per-pixel density), scaled back by 4."""
    from common_bench import oracle_fwd_bwd
    P, W, H, S, seed = CONFIGS[args.config]
    k = 4
    t, cores = oracle_fwd_bwd(max(P // k, 1000), max(W // 2, 64), max(H // 2, 64), S, seed)
    return {"value": round(1.0 / (t * k), 5), "unit": "views/s", "cores": cores, "kind": "port",
            "sample": f"oracle/goi_oracle.c fwd+bwd on P={P // k}, {W // 2}x{H // 2}, S={S} ({t:.2f} s; forward "
                      f"tiles on {cores} OpenMP threads, reverse walk sequential), scaled x1/{k} to the full view"}


def run_reference(args):
    """The reference's own CUDA kernels + glue (oracle/_ref/libref_S*.so) on the same workload, on EVERY rank: N
    independent single-GPU replicas of the reference (it has no multi-GPU path, utils/general_utils.py:144), each
    rendering `views_per_step` views per step with the per-view gradients summed the way its autograd would
    (AccumulateGrad: one add per tensor per view), and a torch NCCL all-reduce of those gradient tensors per step
    (SURVEY.md section 8d)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from oracle import refshim
    from goi_b200.scenes import make_loss_weights
    P, W, H, S, seed = CONFIGS[args.config]
    if not refshim.available(S):
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref/libref_S{S}.so not built"}))
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    g, cams, bg = build_workload(args.config, dev)
    n_cams = len(cams)
    w = {k: v.contiguous() for k, v in make_loss_weights(S, W, H, seed, device=dev).items()}
    rr = refshim.RefRasterizer(S)
    tensors = dict(means3D=g.get_xyz.contiguous(), opacities=g.get_opacity.contiguous(),
                   shs=g.get_features.contiguous(), semantics=g.get_semantics.contiguous(),
                   scales=g.get_scaling.contiguous(), rotations=g.get_rotation.contiguous())
    vps = args.views_per_step
    param_grads = ("dL_dmeans3D", "dL_dopacity", "dL_dscales", "dL_drotations", "dL_dsh", "dL_dsemantics")
    rs = []

    def step(i):
        total = None
        for v in range(vps):
            view = i * vps + v
            cam = cams[(view * world + rank) % n_cams]
            rr.forward(W=W, H=H, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
                       campos=cam.camera_center, tanfovx=math.tan(cam.FoVx / 2), tanfovy=math.tan(cam.FoVy / 2), bg=bg,
                       **tensors)
            gr = rr.backward(w["render"], w["semantics"], w["depth"], w["alpha"])
            if total is None:
                total = {k: gr[k] for k in param_grads}
            else:
                for k in param_grads:
                    total[k] += gr[k]                  # autograd's AccumulateGrad
        if world > 1:
            for k in param_grads:
                dist.all_reduce(total[k], op=dist.ReduceOp.SUM)
        rs.append(rr.num_rendered)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    del rs[:]
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    ms_t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms = float(ms_t.item())
    if rank != 0:
        return
    value = world * vps * args.steps / (ms / 1e3)
    line = {
        "impl": "reference",
        "metric": metric_name(args.config),
        "value": round(value, 3), "unit": "views/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, world, vps),
        "run": run_dict(args, sum(rs) / max(len(rs), 1)),
        "clocks": clocks,
        "ms_per_view": round(ms / args.steps / vps, 4),
        "reference_class": "gpu",
        "cpu_baseline": {"value": round(value, 3), "unit": "views/s", "cores": world, "kind": "reference",
                         "sample": "the reference's own CUDA rasterizer (cuda_rasterizer/*.cu compiled unmodified "
                                   "for sm_100a) driven by one host thread per GPU incl. its blocking num_rendered "
                                   "read-back and zero fills; the reference has no CPU rasterizer"},
        "e2e": {"value": round(value, 3), "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--views-per-step", type=int, default=4,
                    help="training views per GPU per step (gradients summed in place, one all-reduce per step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sync-binning", action="store_true",
                    help="read num_rendered back per view like the reference (default: goi_forward_async, no host sync)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
