"""CPU tier: host logic, the C-ABI surface (load + symbols + argument validation, no compute), workload
shapes, batch sharding, and the world_size-2 gloo path of the multi-GPU accounting."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import codetr_b200 as cb
from codetr_b200 import sharding, workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "msda_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msda_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 10
    lib = ctypes.CDLL(cb._native.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/msda_b200.h but not exported"
    assert sorted(cb._native.EXPORTED_SYMBOLS) == names
    nm = subprocess.run(["nm", "-D", "--defined-only", cb._native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", nm))
    assert set(names) <= exported
    # plain C ABI: no C++-mangled or torch symbols leak out of the boundary library
    assert not [s for s in exported if s.startswith("_Z") and "msda" in s.lower() and "GLOBAL__N" not in s]
    ldd = subprocess.run(["ldd", cb._native.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "c10" not in ldd


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", cb._native.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_abi_version_and_error_strings():
    lib = cb._native.load()
    assert lib.msda_b200_abi_version() == 1
    assert cb._native.error_string(0) == "success"
    for code in (-1, -2, -3, -4, -5, -6, -7):
        assert "msda_b200" in cb._native.error_string(code)
    assert "im2col_step" in cb._native.error_string(-4)


def test_argument_validation_returns_before_any_cuda_work():
    """Rejected calls return a negative code without touching the device (so this runs on CPU)."""
    lib = cb._native.load()
    buf = (ctypes.c_char * 4096)()
    p = ctypes.addressof(buf)
    ok_dims = (1, 16, 2, 4, 1, 3, 2)  # B,S,M,D,L,Q,P
    f = lib.msda_b200_forward
    assert f(p, p, p, p, p, p, *ok_dims, 64, 99, 0, None) == -3                      # unknown dtype
    assert f(p, p, p, p, p, p, -1, 16, 2, 4, 1, 3, 2, 64, 0, 0, None) == -2          # negative dim
    assert f(p, p, p, p, p, None, *ok_dims, 64, 0, 0, None) == -1                    # NULL output
    assert f(None, p, p, p, p, p, *ok_dims, 64, 0, 0, None) == -1                    # NULL value
    assert f(p, p, p, p, p, p, 4, 16, 2, 4, 1, 3, 2, 3, 0, 0, None) == -4            # 4 % min(4,3) != 0
    assert f(p, p, p, p, p, p, 4, 16, 2, 4, 1, 3, 2, 0, 0, 0, None) == -4            # step 0
    assert f(p + 2, p, p, p, p, p, *ok_dims, 64, 0, 0, None) == -5                   # value misaligned for f32
    assert f(p, p, p, p, p, p, *ok_dims, 64, 1, (1 << 2) | (1 << 3), None) == -7     # fhfma and exact together
    assert f(p, p, p, p, p, p, 1, 2 ** 31, 2, 4, 1, 3, 2, 64, 0, 0, None) == -6      # keys beyond the index range
    assert f(p, p, p, p, p, p, 0, 16, 2, 4, 1, 3, 2, 64, 0, 0, None) == 0            # empty batch: nothing to launch
    assert f(p, p, p, p, p, p, 1, 16, 2, 4, 1, 0, 2, 64, 0, 0, None) == 0            # no queries
    assert lib.msda_b200_forward_fused(p, p, p, p, p, p, p, *ok_dims, 3, 0, 0, None) == -2   # ref_dim must be 2 or 4
    vd = (ctypes.c_int64 * 4)(1, 16, 2, 4)
    ld = (ctypes.c_int64 * 6)(1, 3, 2, 1, 2, 2)
    ins = (ctypes.c_void_p * 5)(p, p, p, p, p)
    outs = (ctypes.c_void_p * 1)(p)
    assert lib.msda_b200_plugin_enqueue(vd, ld, 3, ins, outs, None, 0, 64, None) == -3  # kINT8 is not a plugin dtype
    ld_bad = (ctypes.c_int64 * 6)(2, 3, 2, 1, 2, 2)
    assert lib.msda_b200_plugin_enqueue(vd, ld_bad, 0, ins, outs, None, 0, 64, None) == -2
    assert cb.launch_count() == 0 or cb.launch_count() >= 0  # no launch happened in this test
    assert lib.msda_b200_host_workspace_bytes(1, 16, 2, 4, 1, 3, 2, 0) >= (16 * 8 + 3 * 2 * 2 * 3 + 3 * 8) * 4


def test_projection_entry_points_validate_before_any_cuda_work():
    """msda_b200_value_proj / msda_b200_output_proj: what has a tensor-core kernel, and the error codes for everything
    else, without touching the device."""
    lib = cb._native.load()
    sup = lib.msda_b200_value_proj_supported
    F16, BF16, F32 = cb._native.DTYPE_F16, cb._native.DTYPE_BF16, cb._native.DTYPE_F32
    for k, n in ((256, 256), (64, 64), (128, 256), (256, 192), (192, 64)):
        assert sup(k, n, F16) == 1 and sup(k, n, BF16) == 1
    assert sup(256, 256, F32) == 0 and sup(256, 256, 3) == 0                   # 16-bit element types only
    assert sup(320, 256, F16) == 0 and sup(256, 512, F16) == 0                 # wider than the resident weight tile
    assert sup(96, 256, F16) == 0 and sup(256, 96, F16) == 0 and sup(0, 0, F16) == 0   # not a multiple of 64
    buf = (ctypes.c_char * 4096)()
    p = ctypes.addressof(buf)
    vp, op = lib.msda_b200_value_proj, lib.msda_b200_output_proj
    assert vp(p, p, p, p, p, 10, 256, 256, 99, 0, None) == -3                   # unknown dtype
    assert vp(p, p, p, p, p, -1, 256, 256, F16, 0, None) == -2                  # negative row count
    assert vp(p, p, p, p, p, 10, 256, 256, F32, 0, None) == -6                  # no fp32 kernel: caller keeps its GEMM
    assert vp(p, p, p, p, p, 10, 320, 256, F16, 0, None) == -6
    assert vp(p, p, p, p, p, 0, 256, 256, F16, 0, None) == 0                    # no rows: nothing to launch
    assert vp(None, p, p, p, p, 10, 256, 256, F16, 0, None) == -1               # NULL x
    assert vp(p, p, None, None, None, 10, 256, 256, F16, 0, None) == -1         # NULL output (bias / mask may be NULL)
    assert vp(p + 2, p, p, p, p, 10, 256, 256, F16, 0, None) == -6              # TMA needs 16-byte aligned bases
    assert op(p, p, p, None, p, 10, 256, 256, F16, 0, None) == -1               # output_proj needs the residual
    assert op(p, p, p, p, p, 0, 256, 256, BF16, 0, None) == 0


def test_algorithmic_byte_counts_match_survey():
    lib = cb._native.load()
    wl = W.CONFIGS["swinl_enc_1152x768"]
    d = wl.dims()
    # config 3, fp16, B=1 -> 36,533,376 B (+120 B of level tables)
    got = lib.msda_b200_algorithmic_hbm_bytes(1, d["S"], d["M"], d["D"], d["L"], d["Q"], d["P"], 1)
    assert got == 36_533_376 + 120 == W.algorithmic_hbm_bytes(wl, 1, 2)
    gat = lib.msda_b200_algorithmic_gather_bytes(1, d["M"], d["D"], d["L"], d["Q"], d["P"], 1)
    assert gat == 40_960 * 18_414 == W.algorithmic_gather_bytes(wl, 1, 2)
    dec = W.CONFIGS["swinl_dec_1152x768"]
    assert W.algorithmic_hbm_bytes(dec, 1, 2) == 10_752_768 + 120


def test_pyramid_shapes_and_key_counts():
    assert W.pyramid_shapes(768, 1152) == [(96, 144), (48, 72), (24, 36), (12, 18), (6, 9)]
    assert W.CONFIGS["r50_enc_608"].S == 7_706
    assert W.CONFIGS["swinl_enc_1152x768"].S == 18_414
    assert W.CONFIGS["swinl_enc_1920x1280"].S == 51_150
    assert W.CONFIGS["swinl_enc_1152x768_s4"].S == 73_656
    assert W.num_keys(W.pyramid_shapes(1280, 1920, W.STRIDES_REFERENCE)) == 204_600
    assert W.num_keys(W.pyramid_shapes(608, 608, W.STRIDES_REFERENCE)) == 30_785
    assert W.level_starts([(2, 3), (1, 1), (4, 4)]) == [0, 6, 7]


def test_workload_generator_is_deterministic_and_consistent():
    wl = W.Workload(name="t", shapes=tuple(W.pyramid_shapes(64, 96)), num_queries=0, batch=2, kind="encoder", seed=9)
    a, b = W.make_inputs(wl), W.make_inputs(wl)
    for k in ("value", "sampling_loc", "attn_weight", "reference_points", "sampling_offsets", "attn_logits"):
        assert np.array_equal(getattr(a, k), getattr(b, k))
    assert a.value.shape == (2, 129, 8, 32) and a.sampling_loc.shape == (2, 129, 8, 5, 4, 2)
    assert np.allclose(a.attn_weight.reshape(2, 129, 8, -1).sum(-1), 1.0, atol=1e-5)
    # producers reproduce loc: ref + off / (W, H)   (multi_scale_deformable_attention.py:186-191)
    wh = np.stack([a.spatial_shapes[:, 1], a.spatial_shapes[:, 0]], -1).astype(np.float32)
    loc = a.reference_points[:, :, None, :, None, :] + a.sampling_offsets / wh[None, None, None, :, None, :]
    assert np.allclose(loc, a.sampling_loc, atol=1e-6)
    dec = W.make_inputs(wl.with_(kind="decoder", num_queries=10))
    assert dec.reference_points.shape == (2, 10, 5, 4)


def test_fake_kernel_shapes_and_checks():
    """Meta-device call goes through the registered fake kernel (reference: codetr/ops.py:19-87)."""
    mk = lambda *s, dt=torch.float16: torch.empty(*s, dtype=dt, device="meta")
    out = torch.ops.codetr.multi_scale_deformable_attention(
        mk(2, 129, 8, 32), mk(5, 2, dt=torch.int64), mk(5, dt=torch.int64), mk(2, 77, 8, 5, 4, 2), mk(2, 77, 8, 5, 4), 64)
    assert tuple(out.shape) == (2, 77, 256) and out.dtype == torch.float16 and out.device.type == "meta"
    with pytest.raises(RuntimeError):
        torch.ops.codetr.multi_scale_deformable_attention(
            mk(2, 129, 8, 32), mk(5, 2, dt=torch.int32), mk(5, dt=torch.int64), mk(2, 77, 8, 5, 4, 2), mk(2, 77, 8, 5, 4), 64)
    with pytest.raises(RuntimeError):
        torch.ops.codetr.multi_scale_deformable_attention(
            mk(2, 129, 8, 32), mk(5, 2, dt=torch.int64), mk(5, dt=torch.int64), mk(2, 77, 4, 5, 4, 2), mk(2, 77, 8, 5, 4), 64)
    schema = str(torch.ops.codetr.multi_scale_deformable_attention.default._schema)
    assert schema == ("codetr::multi_scale_deformable_attention(Tensor value, Tensor spatial_shapes, Tensor level_start_index, "
                      "Tensor sampling_loc, Tensor attn_weight, int im2col_step) -> Tensor")


def test_cpu_tensors_are_rejected_without_fallback():
    v = torch.zeros(1, 4, 2, 4)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.codetr.multi_scale_deformable_attention(v, torch.tensor([[2, 2]]), torch.tensor([0]),
                                                          torch.zeros(1, 3, 2, 1, 2, 2), torch.zeros(1, 3, 2, 1, 2), 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        cb.multi_scale_deformable_attention(v, torch.tensor([[2, 2]]), torch.tensor([0]), torch.zeros(1, 3, 2, 1, 2, 2),
                                            torch.zeros(1, 3, 2, 1, 2))


def test_image_ranges_partition_the_batch():
    for n in (0, 1, 7, 8, 16, 17):
        for ws in (1, 2, 4, 8):
            rs = sharding.all_ranges(n, ws)
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.image_range(4, 2, 2)
    t = torch.arange(10).reshape(5, 2)
    shapes = torch.tensor([[2, 2]])
    a, s = sharding.shard_batch([t, (shapes, True)], 1, 2)
    assert a.tolist() == [[6, 7], [8, 9]] and s is shapes
    assert sharding.weak_scaling_images(2, 8) == 16


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.image_range(5, rank, world)
    ips, max_ms, total = sharding.aggregate_throughput(10.0 * (rank + 1), hi - lo)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi))
    dist.destroy_process_group()
    q.put((rank, ips, max_ms, total, gathered))


def test_two_rank_gloo_accounting():
    """world_size-2 on CPU: shards cover the batch once, the job time is the slowest rank's, the job
    throughput counts every rank's images."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ips, max_ms, total, gathered in res:
        assert max_ms == 20.0 and total == 5
        assert ips == pytest.approx(5 / 0.020)
        assert gathered == [(0, 3), (3, 5)]


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` needs no GPU: it times the CPU grid_sample path and prints one JSON line."""
    import json

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--workload", "r50_enc_608"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    # the reference arm runs none of this repo's native code: the product library is not even mapped
    assert line["product_so_loaded"] == []
    assert line["config"]["workload"] == "r50_enc_608" and line["config"]["Q"] == 7706 and line["config"]["B"] == 1


def test_missing_native_library_fails_loudly():
    """No CPU / PyTorch fallback: without libmsda_b200.so the package refuses to import."""
    code = "import codetr_b200"
    env = dict(os.environ, MSDA_B200_LIB="/nonexistent/libmsda_b200.so", PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT)
    assert out.returncode != 0
    assert "NativeLibraryError" in out.stderr and "no CPU fallback" in out.stderr


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package, include/ or the C/CUDA sources refers to it."""
    pkg = os.path.join(ROOT, "co-detr-tensorrt_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "msda_oracle" in txt or "oracle/" in txt.replace("``oracle/``", ""):
                    offenders.append(os.path.join(base, f))
    assert not offenders, offenders
    code = "import sys, codetr_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, PYTHONPATH=ROOT), cwd=ROOT)
    assert out.returncode == 0, out.stderr


def test_numa_binding_is_a_safe_no_op_without_a_gpu():
    assert sharding.gpu_numa_cpus(0) is None or isinstance(sharding.gpu_numa_cpus(0), set)
    before = os.sched_getaffinity(0)
    prev = sharding.bind_to_gpu_numa_node(0)
    if prev is not None:
        os.sched_setaffinity(0, prev)
    assert os.sched_getaffinity(0) == before


def test_registered_overload_is_what_the_dynamo_converter_keys_on():
    """The reference's Torch-TensorRT converter is registered on `torch.ops.codetr.multi_scale_deformable_attention.default`
    and reads five tensor inputs plus the `im2col_step` int from the node's args (/root/reference/codetr/ops.py:189-291,
    plugin field built at :253-258): the overload name, arity, argument order / names / types and the single Tensor return
    of the op this package registers must be exactly that."""
    import torch

    import codetr_b200  # noqa: F401  (registers the op)

    packet = torch.ops.codetr.multi_scale_deformable_attention
    assert packet.overloads() == ["default"]
    schema = packet.default._schema
    assert schema.name == "codetr::multi_scale_deformable_attention" and schema.overload_name == ""
    names = [a.name for a in schema.arguments]
    types = [str(a.type) for a in schema.arguments]
    assert names == ["value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight", "im2col_step"]
    assert types == ["Tensor"] * 5 + ["int"]
    assert [str(r.type) for r in schema.returns] == ["Tensor"]
    assert not any(a.alias_info is not None and a.alias_info.is_write for a in schema.arguments)   # functional: safe to trace
    bwd = torch.ops.codetr.multi_scale_deformable_attention_backward.default._schema
    assert [str(a.type) for a in bwd.arguments] == ["Tensor"] * 9 + ["int"]


def test_bench_roofline_helpers_recompute_from_committed_files():
    """Every floor of bench.py's `roofline_detail` must be recomputable from files under profiles/: the committed ncu
    counters of the headline configuration, the gather probe's ceilings, and the live-row count of the inputs."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    c = bench.committed_counters("swinl_enc_1152x768", "float16", 1)
    assert c and c["dram_bytes"] > 20e6 and c["lts_read_sectors"] > 1e6 and c["l1_wavefronts"] > 1e6
    assert c["source"].startswith("profiles/") and os.path.isfile(os.path.join(ROOT, c["source"].split(" ")[0]))
    ceil = bench.gather_probe_ceiling()
    assert ceil and 0.9 < ceil["ldg128_rows_per_clk_per_sm"] < 1.3 and 1.7 < ceil["lds128_rows_per_clk_per_sm"] <= 2.0
    # live corner rows: brute force on a small pyramid against the helper
    wl = W.Workload(name="t", shapes=((6, 9), (3, 5)), num_queries=0, batch=1, kind="encoder", seed=2)
    inp = W.make_inputs(wl, out_of_range_frac=0.2)
    want = 0
    loc = inp.sampling_loc.astype(np.float32)
    for l, (H, Wd) in enumerate(inp.spatial_shapes):
        for x, y in loc[..., l, :, :].reshape(-1, 2):
            xi, yi = np.float32(x) * np.float32(Wd) - np.float32(0.5), np.float32(y) * np.float32(H) - np.float32(0.5)
            if not (xi > -1 and xi < Wd and yi > -1 and yi < H):
                continue
            x0, y0 = int(np.floor(xi)), int(np.floor(yi))
            want += sum(1 for dx in (0, 1) for dy in (0, 1) if 0 <= x0 + dx <= Wd - 1 and 0 <= y0 + dy <= H - 1)
    assert bench.live_corner_rows(inp) == want > 0
    # the reference arm's workload loader does not import the package
    mod = bench.load_workloads_standalone()
    assert mod.CONFIGS["swinl_enc_1152x768"].Q == 18414 and mod.__name__ == "_msda_workloads_standalone"
