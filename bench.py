#!/usr/bin/env python
"""bench.py -- MSDA forward throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A *step* is one pass of the hot path over one per-GPU batch of synthetic Co-DINO inputs (default
workload: BASELINE.json configs[2], the Swin-L encoder shape at 1152x768, strides 8-128, S=Q=18,414,
fp16, one image per GPU per step).  Reported:

  value        images/s over all N GPUs, inputs resident in HBM, CUDA-event time of exactly K steps
               (max over ranks), launches issued back to back through the C ABI
  e2e          the same metric through the host-buffer entry point (pinned host -> device copy of every
               input, kernel, device -> host copy of the result inside the timed region)
  roofline     algorithmic HBM bytes of one launch / measured launch duration vs the measured HBM peak
  roofline_detail  the other bounds of the kernel, every one recomputable from a file under profiles/: measured L2
               read traffic (ncu lts__t_sectors_op_read x 32 B) vs the L2->SM read probe, L1 data-pipe wavefronts
               (ncu) at one per clock per SM, the live corner rows of the call vs the row-gather ceiling measured
               by tools/gather_probe.cu; t_roof = the largest floor, t_roof_frac = t_roof / measured time
  reference_cuda   the reference's own CUDA kernel rebuilt for sm_100a (oracle/_ref, when present) on the same tensors
  cpu_baseline the reference's CPU path (PyTorch grid_sample formulation, oracle/ port) on this box's
               host cores, bounded sample (rank 0, N=1 only)

``--impl reference`` times that CPU path alone (rank 0 only) and prints the same line shape.
Cold-cache policy: the step rotates over enough distinct input sets that their total footprint
exceeds the 126 MB L2 (stated in config.l2_policy).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L2_BYTES = 126 * 1024 * 1024
METRIC = "msda_fwd_images_per_s"
UNIT = "images/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, help="name in codetr_b200.workloads.CONFIGS (default: headline)")
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default: the workload's)")
    ap.add_argument("--dtype", default=None, choices=[None, "float16", "bfloat16", "float32"])
    ap.add_argument("--loc-mode", default=None, choices=[None, "encoder", "decoder", "uniform", "adversarial"])
    ap.add_argument("--flags", type=int, default=None, help="msda_flags bit field (default: library default)")
    ap.add_argument("--api", default="cabi", choices=["cabi", "plugin", "torch_op"],
                    help="how the timed loop calls the library: prepared C-ABI call (default), the TensorRT-enqueue-shaped "
                         "entry (raw pointers, external stream), or torch.ops.codetr.multi_scale_deformable_attention")
    ap.add_argument("--l2-warm", action="store_true", help="reuse ONE input set (inputs stay L2-resident); stated in config.l2_policy")
    ap.add_argument("--workspace", action="store_true", help="give the library a scratch buffer (packed-pyramid path)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cuda-graph", action="store_true", help="replay the K steps from one captured CUDA graph")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the bounded CPU sample")
    ap.add_argument("--no-batch-sweep", action="store_true", help="skip the extra per-GPU batch 2/4/8 rows (N=1 only)")
    ap.add_argument("--no-neighbours", action="store_true", help="skip the extra row for the value_proj tcgen05 kernel (N=1 only)")
    ap.add_argument("--no-fused-row", action="store_true", help="skip the extra row for the producer-fused entry (N=1 only)")
    ap.add_argument("--no-extra-workloads", action="store_true", help="skip the rows for BASELINE configs[3] (decoder) and configs[4] (1920x1280, plugin path)")
    ap.add_argument("--no-reference-cuda", action="store_true", help="skip timing the reference's CUDA kernel (oracle/_ref) beside ours")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic(workload: str, dtype: str, batch: int):
    """dram bytes per launch from the committed `ncu --set full` summary, if there is one."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return d.get(f"{workload}/{dtype}/b{batch}")
    except Exception:
        return None



def load_workloads_standalone():
    """codetr_b200.workloads loaded by file path, WITHOUT importing the package (which dlopens the CUDA library):
    the reference arm must not load any of this repo's native code."""
    import importlib.util

    path = os.path.join(ROOT, "co-detr-tensorrt_b200", "workloads.py")
    spec = importlib.util.spec_from_file_location("_msda_workloads_standalone", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def committed_counters(workload: str, dtype: str, batch: int):
    """Per-launch ncu counters of the committed `ncu --set full` capture of this configuration (profiles/traffic.json):
    {"dram_bytes", "lts_read_sectors", "l1_wavefronts", "kernel", "source"} or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        v = d.get(f"{workload}/{dtype}/b{batch}")
        if isinstance(v, dict):
            return v
        if v is not None:
            return {"dram_bytes": v}
    except Exception:
        pass
    return None


def gather_probe_ceiling():
    """Best measured row-gather rates of one SM (tools/gather_probe.cu, profiles/r02_gather_probe.jsonl):
    64-byte rows per clock through LDG.128 (mode 0) and through conflict-free LDS.128 (mode 3)."""
    path = os.path.join(ROOT, "profiles", "r02_gather_probe.jsonl")
    best = {}
    try:
        with open(path) as f:
            for line in f:
                line = line.strip()
                if not line.startswith("{"):
                    continue
                r = json.loads(line)
                if "mode" in r and r.get("window", 0) in (0, 2160):
                    best[r["mode"]] = max(best.get(r["mode"], 0.0), float(r["rows_per_clk_per_sm"]))
    except Exception:
        return None
    if 0 not in best:
        return None
    return {"ldg128_rows_per_clk_per_sm": best.get(0), "lds128_rows_per_clk_per_sm": best.get(3), "source": "profiles/r02_gather_probe.jsonl"}


def live_corner_rows(inp) -> int:
    """Corner rows the call really fetches: corners inside their level, of samples that pass the reference's
    range test (ms_deform_attn.cu:249, :53-71)."""
    import numpy as np

    loc = inp.sampling_loc.astype(np.float32)
    total = 0
    for l, (H, Wd) in enumerate(inp.spatial_shapes):
        x = loc[..., l, :, 0] * np.float32(Wd) - np.float32(0.5)
        y = loc[..., l, :, 1] * np.float32(H) - np.float32(0.5)
        inside = (x > -1) & (x < Wd) & (y > -1) & (y < H)
        x0, y0 = np.floor(x), np.floor(y)
        for dx in (0, 1):
            for dy in (0, 1):
                total += int((inside & (x0 + dx >= 0) & (x0 + dx <= Wd - 1) & (y0 + dy >= 0) & (y0 + dy <= H - 1)).sum())
    return total


def pcie_probe(dev, piece_bytes, d2h_bytes: int, reps: int, barrier):
    """Host-fabric ceiling of this rank while EVERY rank does the same: the e2e step's host -> device pieces and its
    device -> host result as plain pinned cudaMemcpyAsync calls, no kernel, no library, `reps` steps in steady state,
    timed on the host between barriers like the e2e leg itself.  Two patterns, both reported (GB/s, both directions summed):
      free       uploads on one stream, downloads on another, nothing ties them together
      dependent  the download of step i waits for the upload of step i (the dependency a real call has), 3 output slots
    An earlier version timed 100 queued copies with CUDA events right after the e2e leg; ranks that started late then
    measured a half-idle fabric (27.8 GB/s per rank at N = 8 where the steady state gives 18-21), so the ceiling was too high."""
    import torch

    h_in = [torch.empty(max(n, 1), dtype=torch.uint8).pin_memory().fill_(1) for n in piece_bytes]
    d_in = [torch.empty(max(n, 1), dtype=torch.uint8, device=dev) for n in piece_bytes]
    h_out = [torch.empty(max(d2h_bytes, 1), dtype=torch.uint8).pin_memory() for _ in range(3)]
    d_out = torch.zeros(max(d2h_bytes, 1), dtype=torch.uint8, device=dev)
    s_up, s_down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    total = sum(piece_bytes) + d2h_bytes

    def run(n, dependent):
        for i in range(n):
            with torch.cuda.stream(s_up):
                for dst, src in zip(d_in, h_in):
                    dst.copy_(src, non_blocking=True)
                if dependent:
                    ev = torch.cuda.Event()
                    ev.record(s_up)
            with torch.cuda.stream(s_down):
                if dependent:
                    s_down.wait_event(ev)
                h_out[i % 3].copy_(d_out, non_blocking=True)

    out = {}
    for name, dependent in (("free", False), ("dependent", True)):
        run(3, dependent)
        torch.cuda.synchronize(dev)
        barrier()
        t0 = time.perf_counter()
        run(reps, dependent)
        torch.cuda.synchronize(dev)
        out[name] = total * reps / (time.perf_counter() - t0) / 1e9
        barrier()
    return out


def pin_rank_to_cores(local: int, world: int):
    """Disjoint host-core sets per rank (the launch loops and the pinned-copy set-up of 8 ranks otherwise share
    whatever the scheduler gives them).  Returns the previous affinity, or None when nothing changed."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        if world <= 1 or len(allowed) < 2 * world:
            return None
        per = len(allowed) // world
        mine = allowed[local * per:(local + 1) * per]
        os.sched_setaffinity(0, mine)
        return set(allowed)
    except Exception:
        return None

class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU DURING the timed region with the profiling recipe's
    `nvidia-smi --query-gpu=... -lms` line, run as a separate process (a Python thread would starve behind
    the launch loop's GIL); NVML in-process as a fallback when nvidia-smi is missing."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index: int, period_ms: int = 20):
        import shutil
        import subprocess
        import tempfile

        self.index, self.proc, self.path = index, None, None
        self.samples, self.reasons, self.max_mhz = [], set(), None
        exe = shutil.which("nvidia-smi")
        if exe:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen([exe, "-i", str(index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", str(period_ms)], stdout=self.out, stderr=subprocess.DEVNULL)
        self.t_start = None

    def start(self):
        # drop whatever was sampled while the GPU was idle before the timed region
        self.t_start = time.time()
        if self.proc is not None:
            self.out.flush()
            self.skip = os.path.getsize(self.path)

    def sample_once(self):
        try:  # NVML fallback / extra sample while the queue drains
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.samples.append(int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            for bit, name in ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap")):
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def stop(self):
        if self.proc is None:
            return
        time.sleep(0.03)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        try:
            with open(self.path) as f:
                f.seek(getattr(self, "skip", 0))
                for line in f:
                    parts = [x.strip() for x in line.split(",")]
                    if len(parts) < 6 or not parts[0].isdigit():
                        continue
                    self.samples.append(int(parts[0]))
                    self.max_mhz = int(parts[1]) if parts[1].isdigit() else self.max_mhz
                    for name, val in zip(self.NAMES, parts[2:6]):
                        if val.lower().startswith("active"):
                            self.reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def torch_dtype(name: str):
    import torch

    return {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}[name]


# --------------------------------------------------------------------------------------------
# CPU legs (oracle/ is used here only: cpu_baseline and --impl reference)
# --------------------------------------------------------------------------------------------
def cpu_reference_leg(wl, batch, loc_mode, seconds_budget, steps=None, warmup=1):
    """Times the reference's CPU path (grid_sample formulation) on host cores, fp32, all threads.
    Returns (images_per_s, ms_per_step, cores, sample_description, steps_done)."""
    import torch

    import oracle

    W = sys.modules.get("_msda_workloads_standalone") or load_workloads_standalone()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inp = W.make_inputs(wl, batch=1, loc_mode=loc_mode)
    value = torch.from_numpy(inp.value)
    shapes = torch.from_numpy(inp.spatial_shapes)
    loc_full = torch.from_numpy(inp.sampling_loc)
    w_full = torch.from_numpy(inp.attn_weight)
    Q = loc_full.shape[1]

    def call(q):
        with torch.no_grad():
            return oracle.forward_grid_sample(value, shapes, loc_full[:, :q], w_full[:, :q])

    t0 = time.perf_counter()
    call(Q)
    t_full = time.perf_counter() - t0
    if steps is None:
        steps = max(3, min(30, int(seconds_budget / max(t_full, 1e-3))))
        q_used = Q
    else:
        # the driver chose the step count: bound every step so the whole run fits the budget
        per_step = seconds_budget / max(1, steps + warmup)
        frac = min(1.0, per_step / max(t_full, 1e-6))
        q_used = max(64, int(Q * frac)) if frac < 1.0 else Q
    for _ in range(max(0, warmup - 1)):
        call(q_used)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        call(q_used)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    ms_per_step = 1e3 * total / len(times)
    # one step covers q_used of the Q queries of one image: scale to whole images
    images_per_s = (q_used / Q) * len(times) / total
    sample = (f"{len(times)} calls of multi_scale_deformable_attention_pytorch-equivalent (oracle/grid_sample_port.py) on "
              f"{q_used}/{Q} queries of 1 image of {wl.name}, fp32, torch.set_num_threads({cores})")
    return images_per_s, ms_per_step, cores, sample, len(times)


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return None
    W = load_workloads_standalone()  # not `import codetr_b200`: this arm loads none of the repo's native code

    wl = W.CONFIGS[args.workload or W.HEADLINE]
    batch = args.batch or wl.batch
    budget = 150.0
    ips, ms, cores, sample, done = cpu_reference_leg(wl, batch, args.loc_mode, budget, steps=args.steps, warmup=max(1, min(args.warmup, 3)))
    return {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "note": wl.note, "per_gpu_batch": batch, **{**wl.dims(), "B": batch},
                   "loc_mode": args.loc_mode or wl.kind, "device": "cpu"},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # self-check: shared objects of the PRODUCT (co-detr-tensorrt_b200/) mapped into this process -- must be none
        "product_so_loaded": sorted({ln.split()[-1] for ln in open("/proc/self/maps")
                                     if "co-detr-tensorrt_b200" in ln and ln.rstrip().endswith(".so")}),
    }


# --------------------------------------------------------------------------------------------
# the B200 arm
# --------------------------------------------------------------------------------------------
def time_value_proj(dev, dt, bsz, keys, embed, heads):
    """Device time of msda_b200_value_proj (rows = bsz * keys, K = N = embed) inside a CUDA graph, inputs rotated
    over > L2 worth of buffers, next to what the reference module runs for the same lines (cuBLAS Linear +
    masked_fill).  Roofline against HBM: algorithmic bytes = rows*(K+N)*2 + N*K*2 + rows."""
    import torch
    import torch.nn.functional as F

    import codetr_b200 as cb

    rows = bsz * keys
    n_sets = min(32, max(2, -(-int(1.5 * L2_BYTES) // (rows * 2 * embed * 2))))
    torch.manual_seed(0)
    w = (torch.randn(embed, embed, device=dev) / embed ** 0.5).to(dt)
    b = torch.randn(embed, device=dev).to(dt)
    xs = [torch.randn(bsz, keys, embed, device=dev).to(dt) for _ in range(n_sets)]
    mask = torch.zeros(bsz, keys, dtype=torch.bool, device=dev)
    mask[:, -keys // 10:] = True

    def graphed(fns):
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.inference_mode():
            for f in fns:
                f()
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for f in fns:
                    f()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for _ in range(3):
            s_ev.record()
            for _ in range(20):
                g.replay()
            e_ev.record()
            torch.cuda.synchronize()
            us = 1e3 * s_ev.elapsed_time(e_ev) / (20 * len(fns))
            best = us if best is None else min(best, us)
        return best

    ours = graphed([(lambda x=x: cb.value_proj(x, w, b, mask, num_heads=heads)) for x in xs])
    variant = cb.last_variant()
    lib = graphed([(lambda x=x: F.linear(x, w, b).masked_fill(mask[..., None], 0.0)) for x in xs])
    hbm = rows * 2 * embed * 2 + embed * embed * 2 + rows
    peak, peak_src = measured_peaks()
    return {"kernel": variant, "rows": rows, "K": embed, "N": embed, "us_per_call": ours, "cublas_linear_masked_fill_us": lib,
            "roofline": {"bound": "hbm", "achieved": hbm / ours / 1e3, "peak": peak, "unit": "GB/s", "frac": hbm / ours / 1e3 / peak,
                         "algorithmic_bytes_per_launch": hbm, "peak_source": peak_src},
            "tflops": 2.0 * rows * embed * embed / ours / 1e6,
            "timing": f"CUDA graph of {n_sets} calls on rotating inputs (> L2), device time per call"}


def run_b200(args):
    import numpy as np
    import torch

    import codetr_b200 as cb
    from codetr_b200 import workloads as W

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # pinned buffers allocated below are first-touched next to the GPU when the platform exposes its NUMA node
    # (a no-op on single-node VMs); undone before the CPU-baseline leg, which must see every host core
    previous_affinity = cb.sharding.bind_to_gpu_numa_node(local)
    split_affinity = pin_rank_to_cores(local, world)  # disjoint cores per rank on top of the NUMA binding
    if previous_affinity is None:
        previous_affinity = split_affinity
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    wl = W.CONFIGS[args.workload or W.HEADLINE]
    batch = args.batch or wl.batch
    dtype_name = args.dtype or wl.dtype
    dt = torch_dtype(dtype_name)
    esize = torch.empty((), dtype=dt).element_size()
    dims = wl.dims()
    dims["B"] = batch

    hbm_bytes = W.algorithmic_hbm_bytes(wl, batch, esize)
    gather_bytes = W.algorithmic_gather_bytes(wl, batch, esize)
    n_sets = max(2, -(-int(1.5 * L2_BYTES) // hbm_bytes))
    n_sets = min(n_sets, 64)
    if args.l2_warm:
        n_sets = 1

    # two distinct seeded host sets, uploaded alternately into n_sets distinct device copies
    keys = ("value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight")
    host_sets = []
    live_rows = None
    for i in range(2):
        inp = W.make_inputs(wl, batch=batch, seed=wl.seed + 1000 * rank + i, loc_mode=args.loc_mode)
        if i == 0 and rank == 0:
            live_rows = live_corner_rows(inp)
        hs = {}
        for k in keys:
            t = torch.from_numpy(getattr(inp, k))
            hs[k] = (t if t.dtype == torch.int64 else t.to(dt)).pin_memory()
        host_sets.append(hs)
    calls = []
    dev_sets = []
    for i in range(n_sets):
        hs = host_sets[i % 2]
        d = {k: hs[k].to(dev, non_blocking=True) for k in keys}
        dev_sets.append(d)
        ws = None
        if args.workspace:
            need = cb.workspace_bytes(d["value"], d["sampling_loc"])
            ws = torch.empty(need, dtype=torch.uint8, device=dev) if need else None
        if args.api == "cabi":
            calls.append(cb.PreparedForward(*(d[k] for k in keys), flags=args.flags, workspace=ws))
        elif args.api == "plugin":
            # DeformableAttentionPlugin::enqueue's calling convention: dims from descriptors, raw device pointers,
            # caller-owned output, device-resident int64 shapes, stream passed explicitly
            out_t = torch.empty((batch, dims["Q"], dims["M"] * dims["D"]), dtype=dt, device=dev)
            trt_dt = {torch.float32: cb.ops.TRT_FLOAT, torch.float16: cb.ops.TRT_HALF, torch.bfloat16: cb.ops.TRT_BF16}[dt]
            vd, ld, ptrs = tuple(d["value"].shape), tuple(d["sampling_loc"].shape), [d[k].data_ptr() for k in keys]

            def plugin_call(stream_ptr, _vd=vd, _ld=ld, _ptrs=ptrs, _out=out_t, _keep=d):
                rc = cb.plugin_enqueue(_vd, _ld, trt_dt, _ptrs, _out.data_ptr(), stream_ptr)
                assert rc == 0, rc
            calls.append(plugin_call)
        else:
            def op_call(stream_ptr, _d=d):
                return torch.ops.codetr.multi_scale_deformable_attention(*(_d[k] for k in keys), 64)
            calls.append(op_call)
    torch.cuda.synchronize()

    def barrier():
        if use_dist:
            dist.barrier()

    stream = torch.cuda.current_stream(dev)
    sptr = stream.cuda_stream

    # ---- warm-up ----
    for i in range(max(args.warmup, 3)):
        calls[i % n_sets](sptr)
    torch.cuda.synchronize()
    variant = cb.last_variant()

    graph = None
    if args.cuda_graph:
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(stream)
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for i in range(args.steps):
                    calls[i % n_sets](side.cuda_stream)
        torch.cuda.synchronize()

    # ---- timed region: exactly K steps ----
    sampler = ClockSampler(local)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    launches0 = cb.launch_count()
    sampler.start()
    start.record(stream)
    if graph is not None:
        graph.replay()
    else:
        for i in range(args.steps):
            calls[i % n_sets](sptr)
    end.record(stream)
    sampler.sample_once()  # at least one sample while the queue is still draining
    torch.cuda.synchronize()
    sampler.stop()
    barrier()
    launches = cb.launch_count() - launches0 if graph is None else args.steps
    elapsed_ms = start.elapsed_time(end)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if use_dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms_max = float(t.item())
    ms_per_step = elapsed_ms_max / args.steps
    images_per_s = world * batch * args.steps / (elapsed_ms_max * 1e-3)

    # ---- per-call distribution (median and min of individually timed calls) ----
    n_ind = min(200, args.steps)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_ind)]
    for i, (e0, e1) in enumerate(evs):
        e0.record(stream)
        calls[i % n_sets](sptr)
        e1.record(stream)
    torch.cuda.synchronize()
    per_call = sorted(1e3 * e0.elapsed_time(e1) for e0, e1 in evs)
    per_call_us = {"median": per_call[len(per_call) // 2], "min": per_call[0], "p90": per_call[int(0.9 * (len(per_call) - 1))],
                   "calls": n_ind, "note": "each call bracketed by its own event pair (includes event overhead)"}

    # ---- end-to-end leg: host buffers through msda_b200_forward_host ----
    e2e = None
    if not args.no_e2e:
        hs = host_sets[0]
        h2d, d2h = cb.HostForward.bytes_moved(*(hs[k] for k in keys))
        e2e_steps = max(3, min(args.steps, 400))
        # (a) one call at a time, synchronised after each (latency view)
        hf = cb.HostForward(dev)
        out_host = torch.empty((batch, dims["Q"], dims["M"] * dims["D"]), dtype=dt).pin_memory()
        for _ in range(3):
            hf(*(hs[k] for k in keys), output=out_host)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(min(e2e_steps, 100)):
            hf(*(hs[k] for k in keys), output=out_host)
        t_sync = (time.perf_counter() - t0) / min(e2e_steps, 100)
        # (b) the throughput API: 3-deep software pipeline, every step still copies all of its inputs
        # host->device and its result device->host inside the timed region
        pipe = cb.HostPipeline(dev, depth=3)
        for _ in range(6):
            pipe.submit(*(hs[k] for k in keys))
        pipe.drain()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        checksum = 0.0
        for i in range(e2e_steps):
            ticket = pipe.submit(*(hs[k] for k in keys))
            if i % 64 == 63:
                checksum += float(pipe.result(ticket)[0, 0, 0])  # read a result on the host while the pipeline runs
        pipe.drain()
        t_e2e = time.perf_counter() - t0
        my_t = t_e2e
        te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        per_rank_ms = [1e3 * t_e2e / e2e_steps]
        if use_dist:
            allt = [torch.zeros_like(te) for _ in range(world)]
            dist.all_gather(allt, te)
            per_rank_ms = [1e3 * float(x.item()) / e2e_steps for x in allt]
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        t_e2e = float(te.item())
        # host-fabric ceiling, measured the same way on every rank AT THE SAME TIME (inside one barrier window): the
        # step's H2D and D2H byte counts as plain pinned copies, both directions in flight -- no kernel, no library
        pipe = None
        pieces = [hs[k].numel() * hs[k].element_size() for k in keys]
        probe = pcie_probe(dev, pieces, d2h, e2e_steps, barrier)
        tp = torch.tensor([probe["free"], probe["dependent"]], dtype=torch.float64, device=dev)
        per_rank_probe = [[probe["free"], probe["dependent"]]]
        if use_dist:
            allp = [torch.zeros_like(tp) for _ in range(world)]
            dist.all_gather(allp, tp)
            per_rank_probe = [[float(x[0].item()), float(x[1].item())] for x in allp]
        achieved_gbps = (h2d + d2h) * e2e_steps / t_e2e / 1e9  # of the slowest rank
        slowest_peak = min(max(p) for p in per_rank_probe)      # the better pattern, on the rank that gets least
        e2e = {"value": world * batch * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": 1e3 * t_e2e / e2e_steps,
               "pcie_GBps": achieved_gbps,
               "pcie_peak_GBps": slowest_peak, "frac_of_pcie_peak": achieved_gbps / slowest_peak,
               "pcie_peak_note": "the step's byte counts as plain pinned cudaMemcpyAsync calls (no kernel, no library), all ranks at "
                                 "once, same number of steps, host-timed between barriers like the e2e leg; per rank [free, dependent] "
                                 "= [uploads and downloads on independent streams, download i waits for upload i]; the ceiling is the "
                                 "better pattern on the rank that gets least (the max-over-ranks time is set by the slowest rank)",
               "per_rank_ms_per_step": per_rank_ms, "per_rank_pcie_peak_GBps": per_rank_probe,
               "aggregate_GBps": sum((h2d + d2h) / (ms * 1e-3) / 1e9 for ms in per_rank_ms),
               "aggregate_pcie_peak_GBps": sum(max(p) for p in per_rank_probe),
               "host_cores_per_rank": len(os.sched_getaffinity(0)),
               "synchronous_ms_per_call": 1e3 * t_sync,
               "api": "codetr_b200.HostPipeline(depth=3) -> msda_b200_forward_host (pinned host buffers, "
                      "H2D of all inputs + kernel + D2H of the result every step)"}

    # ---- extra rows (N=1 only): the same workload at larger per-GPU batches, one launch per batch ----
    batch_sweep = None
    if world == 1 and not args.no_batch_sweep:
        batch_sweep = {}
        for bsz in (2, 4, 8):
            if bsz == batch:
                continue
            inp_b = W.make_inputs(wl, batch=bsz, seed=wl.seed + 77, loc_mode=args.loc_mode)
            sets_b = []
            n_b = max(2, -(-int(1.5 * L2_BYTES) // W.algorithmic_hbm_bytes(wl, bsz, esize)))
            for i in range(n_b):
                d = {}
                for k in keys:
                    t = torch.from_numpy(getattr(inp_b, k))
                    d[k] = t.to(dev) if t.dtype == torch.int64 else t.to(device=dev, dtype=dt)
                sets_b.append(cb.PreparedForward(*(d[k] for k in keys), flags=args.flags))
            for i in range(5):
                sets_b[i % n_b](sptr)
            torch.cuda.synchronize()
            iters = max(20, min(args.steps, 400) // bsz)
            s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_ev.record(stream)
            for i in range(iters):
                sets_b[i % n_b](sptr)
            e_ev.record(stream)
            torch.cuda.synchronize()
            us = 1e3 * s_ev.elapsed_time(e_ev) / iters
            batch_sweep[f"b{bsz}"] = {"us_per_call": us, "images_per_s": bsz / (us * 1e-6),
                                      "gather_GBps": W.algorithmic_gather_bytes(wl, bsz, esize) / us / 1e3}
            del sets_b
            torch.cuda.empty_cache()

    # ---- extra row (N=1): the opt-in producer-fused entry (softmax + sampling-location arithmetic in-kernel,
    # SURVEY 8(f).1) under the same protocol as the headline: rotating cold input sets, CUDA events, its own HBM roofline.
    # Algorithmic bytes: value + reference points + offsets + logits read once, output written once.
    fused_row = None
    if world == 1 and not args.no_fused_row and wl.kind in ("encoder", "decoder"):
        try:
            lib = cb._native.load()
            fsets = []
            for i in range(n_sets):
                finp = W.make_inputs(wl, batch=batch, seed=wl.seed + i % 2)
                cast = lambda a: torch.from_numpy(a).to(device=dev, dtype=dt)
                fd = {"value": cast(finp.value), "ref": cast(finp.reference_points), "off": cast(finp.sampling_offsets),
                      "logits": cast(finp.attn_logits), "shapes": torch.from_numpy(finp.spatial_shapes).to(dev),
                      "starts": torch.from_numpy(finp.level_start_index).to(dev),
                      "out": torch.empty((batch, dims["Q"], dims["M"] * dims["D"]), dtype=dt, device=dev)}
                fsets.append(fd)
            ref_dim = int(fsets[0]["ref"].shape[-1])

            def fused_call(fd):
                rc = lib.msda_b200_forward_fused(fd["value"].data_ptr(), fd["shapes"].data_ptr(), fd["starts"].data_ptr(), fd["ref"].data_ptr(),
                                                 fd["off"].data_ptr(), fd["logits"].data_ptr(), fd["out"].data_ptr(), batch, dims["S"], dims["M"],
                                                 dims["D"], dims["L"], dims["Q"], dims["P"], ref_dim, cb.ops._DTYPES[dt], 0, sptr)
                assert rc == 0, rc
            for i in range(5):
                fused_call(fsets[i % n_sets])
            torch.cuda.synchronize()
            f_steps = max(20, min(args.steps, 400))
            s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_ev.record(stream)
            for i in range(f_steps):
                fused_call(fsets[i % n_sets])
            e_ev.record(stream)
            torch.cuda.synchronize()
            f_us = 1e3 * s_ev.elapsed_time(e_ev) / f_steps
            f_bytes = esize * batch * (dims["S"] * dims["M"] * dims["D"] + dims["Q"] * dims["L"] * ref_dim
                                       + 3 * dims["Q"] * dims["M"] * dims["L"] * dims["P"] + dims["Q"] * dims["M"] * dims["D"]) + 24 * dims["L"]
            peak_f, _ = measured_peaks()
            fused_row = {"us_per_call": f_us, "images_per_s": batch / (f_us * 1e-6), "kernel": cb.last_variant(), "steps": f_steps,
                         "roofline": {"bound": "hbm", "achieved": f_bytes / f_us / 1e3, "peak": peak_f, "unit": "GB/s",
                                      "frac": f_bytes / f_us / 1e3 / peak_f, "algorithmic_bytes_per_launch": f_bytes},
                         "note": "replaces the caller's softmax + location kernels (17.7 MB written and re-read per image at the headline "
                                 "shape) at the price of a slower sampling kernel; same rotating cold-input protocol as `value`"}
            del fsets
            torch.cuda.empty_cache()
        except Exception as exc:  # pragma: no cover
            fused_row = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.synchronize()

    # ---- extra row (N=1, 16-bit only): the producer of `value` (Linear + masked_fill, tcgen05 kernel) ----
    neighbours = None
    if world == 1 and not args.no_neighbours and esize == 2:
        neighbours = {}
        embed = wl.num_heads * wl.channels
        for bsz in (batch, 4 * batch):
            try:  # an extra row must never cost the headline line
                neighbours[f"value_proj_b{bsz}"] = time_value_proj(dev, dt, bsz, wl.S, embed, wl.num_heads)
            except Exception as exc:  # pragma: no cover
                neighbours[f"value_proj_b{bsz}"] = {"error": f"{type(exc).__name__}: {exc}"}
                torch.cuda.synchronize()


    # ---- the reference's own CUDA kernel (ms_deform_attn.cu:211-261), rebuilt for sm_100a under oracle/_ref, on the
    # same rotating tensors: "the kernel to beat" in the same run (rank 0's GPU, fp16 / fp32 only) ----
    reference_cuda = None
    ref_so = os.path.join(ROOT, "oracle", "_ref", "msda_ref_cuda.so")
    if rank == 0 and not args.no_reference_cuda and dtype_name in ("float16", "float32") and os.path.isfile(ref_so):
        try:
            torch.ops.load_library(ref_so)
            ref_fns = [(lambda d=d: torch.ops.codetr_ref.msda_forward(*(d[k] for k in keys), 64)) for d in dev_sets]
            for i in range(5):
                ref_fns[i % n_sets]()
            torch.cuda.synchronize()
            r_steps = max(10, min(args.steps, 100))
            s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_ev.record(stream)
            for i in range(r_steps):
                ref_fns[i % n_sets]()
            e_ev.record(stream)
            torch.cuda.synchronize()
            ref_us = 1e3 * s_ev.elapsed_time(e_ev) / r_steps
            ours = calls[0](sptr) if args.api == "cabi" else None
            theirs = ref_fns[0]()
            torch.cuda.synchronize()
            reference_cuda = {"us_per_call": ref_us, "images_per_s": batch / (ref_us * 1e-6), "steps": r_steps,
                              "speedup": ref_us / (ms_per_step * 1e3),
                              "kernel": "codetr_ref::msda_forward = ms_deformable_im2col_gpu_kernel (+ the reference's two memsets), "
                                        "oracle/_ref/msda_ref_cuda.so built from /root/reference/codetr/csrc/ms_deform_attn.cu"}
            if ours is not None:
                reference_cuda["max_abs_diff_vs_ours"] = float((ours.float() - theirs.float()).abs().max())
                reference_cuda["max_abs_ref"] = float(theirs.float().abs().max())
        except Exception as exc:  # an extra row must never cost the headline line
            reference_cuda = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.synchronize()

    # ---- the other BASELINE configs on every rank (weak scaling, max over ranks): configs[3] decoder cross-attention
    # (B=1 per GPU: prepared C-ABI call, CUDA-graph replay, and the registered torch op so its host cost shows) and
    # configs[4] 1920x1280 (B=2 per GPU through the TensorRT-enqueue-shaped entry) ----
    extra_workloads = None
    if not args.no_extra_workloads and args.workload is None:
        extra_workloads = {}
        peak_hbm, _ = measured_peaks()

        def timed(fn_list, steps_x, use_graph=False):
            for i in range(5):
                fn_list[i % len(fn_list)](sptr)
            torch.cuda.synchronize()
            g = None
            if use_graph:
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(stream)
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g, stream=side):
                        for i in range(steps_x):
                            fn_list[i % len(fn_list)](side.cuda_stream)
                torch.cuda.synchronize()
            s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s_ev.record(stream)
            if g is not None:
                g.replay()
            else:
                for i in range(steps_x):
                    fn_list[i % len(fn_list)](sptr)
            e_ev.record(stream)
            host_s = time.perf_counter() - t0
            torch.cuda.synchronize()
            ms = torch.tensor([s_ev.elapsed_time(e_ev)], dtype=torch.float64, device=dev)
            if use_dist:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item()) / steps_x, 1e6 * host_s / steps_x

        # ... and the headline shape in the other two element types the plugin accepts (bf16, fp32), same kernel family
        for xname, xbatch, xapis, xdtype in (("swinl_dec_1152x768", 1, ("cabi", "cuda_graph", "torch_op"), None),
                                             ("swinl_enc_1920x1280", 2, ("plugin",), None),
                                             ("swinl_enc_1152x768", 1, ("cabi",), "bfloat16"), ("swinl_enc_1152x768", 1, ("cabi",), "float32")):
            xkey = xname if xdtype is None else f"{xname}/{xdtype}"
            try:
                xwl = W.CONFIGS[xname]
                xdt = torch_dtype(xdtype or xwl.dtype)
                xes = torch.empty((), dtype=xdt).element_size()
                xhbm = W.algorithmic_hbm_bytes(xwl, xbatch, xes)
                xn = min(16, max(2, -(-int(1.5 * L2_BYTES) // xhbm)))
                xinp = W.make_inputs(xwl, batch=xbatch, seed=xwl.seed + 1000 * rank)
                xsets = []
                for i in range(xn):
                    xd = {}
                    for k in keys:
                        t = torch.from_numpy(getattr(xinp, k))
                        xd[k] = t.to(dev) if t.dtype == torch.int64 else t.to(device=dev, dtype=xdt)
                    xsets.append(xd)
                xdims = {**xwl.dims(), "B": xbatch}
                row = {"per_gpu_batch": xbatch, **xdims, "dtype": xdtype or xwl.dtype, "note": xwl.note,
                       "l2_policy": f"rotating {xn} distinct device input sets ({xn * xhbm / 1e6:.0f} MB)", "apis": {}}
                xsteps = max(50, min(args.steps, 400))
                for api in xapis:
                    if api in ("cabi", "cuda_graph"):
                        fns = [cb.PreparedForward(*(xd[k] for k in keys)) for xd in xsets]
                    elif api == "plugin":
                        trt_dt = {torch.float32: cb.ops.TRT_FLOAT, torch.float16: cb.ops.TRT_HALF, torch.bfloat16: cb.ops.TRT_BF16}[xdt]
                        fns = []
                        for xd in xsets:
                            out_t = torch.empty((xbatch, xdims["Q"], xdims["M"] * xdims["D"]), dtype=xdt, device=dev)
                            vd, ld, ptrs = tuple(xd["value"].shape), tuple(xd["sampling_loc"].shape), [xd[k].data_ptr() for k in keys]

                            def plugin_call(stream_ptr, _vd=vd, _ld=ld, _ptrs=ptrs, _out=out_t, _keep=xd):
                                rc = cb.plugin_enqueue(_vd, _ld, trt_dt, _ptrs, _out.data_ptr(), stream_ptr)
                                assert rc == 0, rc
                            fns.append(plugin_call)
                    else:
                        fns = [(lambda stream_ptr, _d=xd: torch.ops.codetr.multi_scale_deformable_attention(*(_d[k] for k in keys), 64))
                               for xd in xsets]
                    ms, host_us = timed(fns, xsteps, use_graph=(api == "cuda_graph"))
                    row["apis"][api] = {"us_per_call": ms * 1e3, "images_per_s": world * xbatch / (ms * 1e-3),
                                        "host_us_per_call": None if api == "cuda_graph" else host_us,
                                        "hbm_frac": xhbm / (ms * 1e-3) / 1e9 / peak_hbm, "kernel": cb.last_variant(), "steps": xsteps}
                extra_workloads[xkey] = row
                del xsets
                torch.cuda.empty_cache()
            except Exception as exc:  # pragma: no cover
                extra_workloads[xkey] = {"error": f"{type(exc).__name__}: {exc}"}
                torch.cuda.synchronize()

    # L2 -> SM read bandwidth of this GPU, measured live with the library's read probe (48 MB working set, L2-resident)
    l2_peak = cb.read_bandwidth_probe(dev, 48 * 1024 * 1024, 24) if rank == 0 else None

    if use_dist:
        dist.destroy_process_group()
    if rank != 0:
        return None

    # ---- roofline of the (single) kernel of the step ----
    peak, peak_src = measured_peaks()
    launch_s = ms_per_step * 1e-3  # one launch per step, back to back on one stream
    achieved = hbm_bytes / launch_s / 1e9
    counters = committed_counters(wl.name, dtype_name, batch)
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": (counters or {}).get("dram_bytes"), "peak_source": peak_src,
        "algorithmic_bytes_per_launch": hbm_bytes, "kernel": variant, "launch_us": launch_s * 1e6,
        "launch_us_note": "back-to-back launches with programmatic dependent launch: a pipelined figure; the same kernel "
                          "timed alone is per_call_us.median (own event pair per call)",
        "isolated_launch_us": per_call_us["median"],
    }

    # ---- the kernel's other bounds (T_roof = max of the floors) ----
    clocks = sampler.summary()
    sm_hz = 1e6 * float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    floors = {"hbm": {"floor_us": hbm_bytes / (peak * 1e9) * 1e6, "bytes": hbm_bytes, "peak_GBps": peak, "source": "algorithmic bytes / " + peak_src}}
    if counters and counters.get("lts_read_sectors"):
        l2_bytes = 32 * int(counters["lts_read_sectors"])
        floors["l2_actual"] = {"floor_us": l2_bytes / (l2_peak * 1e9) * 1e6, "bytes": l2_bytes, "peak_GBps": l2_peak,
                               "achieved_GBps": l2_bytes / launch_s / 1e9, "frac_of_l2_peak": l2_bytes / launch_s / 1e9 / l2_peak,
                               "source": "ncu lts__t_sectors_op_read.sum x 32 B (" + str(counters.get("source")) + ") / msda_b200_read_probe, 48 MB, this run"}
    if counters and counters.get("l1_wavefronts"):
        wf = int(counters["l1_wavefronts"])
        floors["l1_wavefront"] = {"floor_us": wf / (sms * sm_hz) * 1e6, "wavefronts": wf, "sm_clock_MHz": sm_hz / 1e6,
                                  "source": "ncu l1tex__data_pipe_lsu_wavefronts.sum (" + str(counters.get("source")) + ") at 1 / clk / SM"}
    ceiling = gather_probe_ceiling()
    if ceiling and live_rows:
        rate = float(ceiling["ldg128_rows_per_clk_per_sm"])
        floors["row_gather"] = {"floor_us": live_rows / (rate * sms * sm_hz) * 1e6, "live_corner_rows": live_rows,
                                "rows_per_clk_per_sm": rate, "sm_clock_MHz": sm_hz / 1e6,
                                "source": "64-byte corner rows of the call's first input set / LDG.128 row-gather ceiling of " + ceiling["source"]}
    t_roof_name = max(floors, key=lambda k: floors[k]["floor_us"])
    roofline_detail = {
        "floors": floors, "t_roof_us": floors[t_roof_name]["floor_us"], "t_roof_bound": t_roof_name,
        "t_roof_frac": floors[t_roof_name]["floor_us"] / (launch_s * 1e6),
        "t_roof_frac_isolated": floors[t_roof_name]["floor_us"] / per_call_us["median"],
        "no_reuse_gather_bytes": gather_bytes,
        "note": "frac = floor / measured launch time; ncu counters come from the committed capture of this configuration "
                "(profiles/traffic.json), peaks and time from this run",
    }

    if previous_affinity:
        os.sched_setaffinity(0, previous_affinity)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ips, ms, cores, sample, _ = cpu_reference_leg(wl, batch, args.loc_mode, args.cpu_seconds)
        cpu = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "ms_per_image": ms}

    return {
        "metric": METRIC, "value": images_per_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "us_per_call": ms_per_step * 1e3,
        "per_call_us": per_call_us,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"float16": "f16", "bfloat16": "bf16", "float32": "f32"}[dtype_name], "data": "synthetic",
        "config": {
            "workload": wl.name, "note": wl.note, "per_gpu_batch": batch, **dims,
            "loc_mode": args.loc_mode or wl.kind, "sharding": f"batch by image, {world} rank(s), no collective on the data path",
            "l2_policy": (f"rotating {n_sets} distinct input sets, {n_sets * hbm_bytes / 1e6:.0f} MB > 126 MB L2" if n_sets > 1
                          else "L2-WARM: one input set reused every step (not a cold-cache number)"),
            "launch": "cuda_graph" if graph is not None else {"cabi": "C ABI via ctypes, back to back on one stream (launches carry the programmatic-stream-serialization attribute; reads wait for the previous kernel)",
                      "plugin": "msda_b200_plugin_enqueue (TensorRT enqueue convention), back to back on one stream",
                      "torch_op": "torch.ops.codetr.multi_scale_deformable_attention, back to back"}[args.api],
        },
        "roofline": roofline, "roofline_detail": roofline_detail, "reference_cuda": reference_cuda,
        "extra_workloads": extra_workloads, "fused_producers": fused_row, "batch_sweep": batch_sweep, "neighbour_kernels": neighbours, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks,
    }


def main():
    args = parse_args()
    # Exactly ONE line may reach stdout.  Libraries chat there (NCCL prints "NCCL version ..." on init), so
    # stdout is pointed at stderr for the duration of the run and the JSON line goes to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run_reference(args) if args.impl == "reference" else run_b200(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
    if line is not None:
        os.write(1, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
