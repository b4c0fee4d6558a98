"""Where does the host-buffer path lose time against the raw PCIe copies?  (manual, GPU box, any N)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/e2e_probe.py
Every variant moves the headline step's bytes (27.1 MB in, 9.4 MB out) per step on every rank at once; reported: ms per
step of the slowest rank and of each rank.
  raw1      one 27.1 MB H2D + one 9.4 MB D2H per step, two streams, all steps queued up front (bench.py's ceiling probe)
  raw5      the H2D split into the call's five pieces (value, shapes, starts, locations, weights)
  raw5dep   like raw5, and the D2H of step i waits for the H2D of step i (the dependency the real call has), 3 slots
  pipeD     codetr_b200.HostPipeline(depth=D): the real call (copies + kernel + copy back)
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import codetr_b200 as cb
from codetr_b200 import workloads as W

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
try:
    allowed = sorted(os.sched_getaffinity(0)); per = len(allowed) // world
    if world > 1 and per >= 2: os.sched_setaffinity(0, allowed[local * per:(local + 1) * per])
except Exception: pass
wl = W.CONFIGS[W.HEADLINE]
inp = W.make_inputs(wl, batch=1, seed=wl.seed + rank)
keys = ("value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight")
hs = {}
for k in keys:
    t = torch.from_numpy(getattr(inp, k))
    hs[k] = (t if t.dtype == torch.int64 else t.half()).pin_memory()
h2d, d2h = cb.HostForward.bytes_moved(*(hs[k] for k in keys))
STEPS = int(os.environ.get("STEPS", 150))

def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

def report(name, seconds):
    t = torch.tensor([seconds], dtype=torch.float64, device=dev)
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per = [1e3 * float(x) / STEPS for x in allt]
    else:
        per = [1e3 * seconds / STEPS]
    if rank == 0:
        print(f"{name:10s} slowest {max(per):7.3f} ms/step = {(h2d + d2h) / max(per) / 1e6:6.1f} GB/s per rank | per rank " + " ".join(f"{p:.3f}" for p in per), flush=True)

# ---- raw copies ----
pieces = [hs[k].view(torch.uint8).reshape(-1) if hs[k].dtype != torch.int64 else hs[k].view(torch.uint8).reshape(-1) for k in keys]
one = torch.empty(h2d, dtype=torch.uint8).pin_memory()
d_one = torch.empty(h2d, dtype=torch.uint8, device=dev)
d_pieces = [torch.empty(p.numel(), dtype=torch.uint8, device=dev) for p in pieces]
h_out = [torch.empty(d2h, dtype=torch.uint8).pin_memory() for _ in range(3)]
d_out = torch.zeros(d2h, dtype=torch.uint8, device=dev)
s_up, s_down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

def raw(split, dep):
    def run(n):
        for i in range(n):
            with torch.cuda.stream(s_up):
                if split:
                    for dp, p in zip(d_pieces, pieces): dp.copy_(p, non_blocking=True)
                else:
                    d_one.copy_(one, non_blocking=True)
                ev = torch.cuda.Event(); ev.record(s_up)
            with torch.cuda.stream(s_down):
                if dep: s_down.wait_event(ev)
                h_out[i % 3].copy_(d_out, non_blocking=True)
    run(3); barrier()
    t0 = time.perf_counter(); run(STEPS); torch.cuda.synchronize(); return time.perf_counter() - t0

for name, split, dep in (("raw1", False, False), ("raw5", True, False), ("raw5dep", True, True), ("raw1", False, False)):
    barrier(); report(name, raw(split, dep))

# ---- the real call ----
for depth in (2, 3, 4, 6):
    pipe = cb.HostPipeline(dev, depth=depth)
    for _ in range(2 * depth): pipe.submit(*(hs[k] for k in keys))
    pipe.drain(); barrier()
    t0 = time.perf_counter()
    for i in range(STEPS): pipe.submit(*(hs[k] for k in keys))
    pipe.drain(); dt = time.perf_counter() - t0
    barrier(); report(f"pipe{depth}", dt)
if world > 1: dist.destroy_process_group()
