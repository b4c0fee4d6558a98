"""H2D / D2H bandwidth of pinned vs write-combined pinned host memory (manual, GPU box)."""
import ctypes, os, sys, time, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import codetr_b200
print("cpus allowed", len(os.sched_getaffinity(0)), "gpu numa cpus", codetr_b200.sharding.gpu_numa_cpus(0) and len(codetr_b200.sharding.gpu_numa_cpus(0)))
os.system("cat /sys/devices/system/node/node*/cpulist 2>/dev/null | head -4; nproc")
if len(sys.argv) > 1 and sys.argv[1] == "bind": print("bound, previous:", len(codetr_b200.sharding.bind_to_gpu_numa_node(0) or []))
rt = torch.cuda.cudart()
dev = torch.device("cuda:0")
n = 27 * 1024 * 1024
d = torch.empty(n, dtype=torch.uint8, device=dev)

_cudart = ctypes.CDLL("libcudart.so.12")
_cudart.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]

def wc_tensor(nbytes, flags):
    ptr = ctypes.c_void_p()
    err = _cudart.cudaHostAlloc(ctypes.byref(ptr), nbytes, flags)
    assert int(err) == 0, err
    buf = (ctypes.c_char * nbytes).from_address(int(ptr.value))
    return torch.frombuffer(buf, dtype=torch.uint8), ptr

def bw(src, dst, reps=50):
    for _ in range(5): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): dst.copy_(src, non_blocking=True)
    e.record(); torch.cuda.synchronize()
    return src.numel() * reps / (s.elapsed_time(e) * 1e-3) / 1e9

plain = torch.empty(n, dtype=torch.uint8).pin_memory()
print("H2D pinned          %.1f GB/s" % bw(plain, d))
print("D2H pinned          %.1f GB/s" % bw(d, plain))
for flags, name in ((4, "write-combined"), (0, "cudaHostAlloc default"), (1, "portable")):
    try:
        t, _ = wc_tensor(n, flags)
        t.fill_(3)
        # torch does not know this memory is pinned: use cudaMemcpyAsync directly
        stream = torch.cuda.current_stream().cuda_stream
        lib = _cudart
        lib.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        def run(reps):
            for _ in range(reps): lib.cudaMemcpyAsync(d.data_ptr(), t.data_ptr(), n, 1, stream)
        run(5); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(50); e.record(); torch.cuda.synchronize()
        print("H2D %-22s %.1f GB/s" % (name, n * 50 / (s.elapsed_time(e) * 1e-3) / 1e9))
    except Exception as ex:
        print(name, "failed", ex)
