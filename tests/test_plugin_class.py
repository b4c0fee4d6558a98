"""The libtorch-free TensorRT plugin class (co-detr-tensorrt_b200/csrc/deformable_attention_plugin_b200.cpp).

TensorRT is not in this image, so the file is compiled against the API-shaped stub headers under tests/stubs/ and
driven through the virtual calls TensorRT makes by tests/stubs/plugin_harness.cpp.  CPU tier: it compiles, keeps the
reference plugin's identity and serialisation layout (/root/reference/codetr/csrc/deformable_attention_plugin.cpp:77-79,
:84-86, :381-388, :398), accepts fp32 / fp16 / bf16 (:218-246 + kBF16), reports shapes / dtypes / workspace like the
reference (:248-281, :371), and never aborts on a malformed network.  GPU tier: `enqueue` on raw device pointers and
an external stream equals the registered op bit for bit.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(ROOT, "tests", "stubs")
CSRC = os.path.join(ROOT, "co-detr-tensorrt_b200", "csrc")
PLUGIN_SRC = os.path.join(CSRC, "deformable_attention_plugin_b200.cpp")
OUT = os.path.join(STUBS, "_build", "libdeformable_attention_plugin_b200_stub.so")

TRT_FLOAT, TRT_HALF, TRT_INT8, TRT_INT32, TRT_BF16, TRT_INT64 = 0, 1, 2, 3, 7, 8
FIELD_INT64, FIELD_UNKNOWN = 10, 8
LINEAR, CHW4 = 0, 3


@pytest.fixture(scope="module")
def harness():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = [PLUGIN_SRC, os.path.join(STUBS, "plugin_harness.cpp")]
    if not os.path.isfile(OUT) or any(os.path.getmtime(s) > os.path.getmtime(OUT) for s in srcs + [os.path.join(STUBS, "NvInfer.h")]):
        cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-Wall", "-Wextra", "-Werror", "-I", STUBS, "-I", os.path.join(ROOT, "include"),
               *srcs, "-L", CSRC, "-lmsda_b200", f"-Wl,-rpath,{CSRC}", "-o", OUT]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        assert proc.returncode == 0, proc.stderr
    lib = ctypes.CDLL(OUT)
    lib.harness_last_error.restype = ctypes.c_char_p
    return lib


def _text(fn, *args):
    buf = ctypes.create_string_buffer(256)
    assert fn(buf, 256, *args) == 0
    return buf.value.decode()


def test_plugin_source_has_no_libtorch_dependency(harness):
    src = open(PLUGIN_SRC).read()
    code = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("//"))
    assert "ATen" not in code and "c10" not in code and "torch" not in code
    ldd = subprocess.run(["ldd", OUT], capture_output=True, text=True).stdout
    assert "libtorch" not in ldd and "libc10" not in ldd and "libmsda_b200.so" in ldd
    # the two C entry points TensorRT's plugin loader resolves (reference plugin.cpp:507-514)
    assert hasattr(harness, "getPluginCreators") and hasattr(harness, "setLoggerFinder")
    assert harness.harness_registered_creators() == 1            # REGISTER_TENSORRT_PLUGIN ran at load time (:466)


def test_identity_and_creator_fields_match_the_reference(harness):
    assert _text(harness.harness_creator_identity) == f"DeformableAttentionPlugin|1||im2col_step|{FIELD_INT64}|1|1"
    assert harness.harness_create_build(32) == 0
    assert _text(harness.harness_core_identity) == "DeformableAttentionPlugin|1|"


def test_serialised_state_is_the_reference_layout_and_round_trips(harness):
    assert harness.harness_create_build(32) == 0
    raw = (ctypes.c_ubyte * 16)()
    assert _text(harness.harness_serialised, raw, 16) == f"parameters|{FIELD_UNKNOWN}|8|1"     # plugin.cpp:381-388
    assert bytes(raw[:8]) == np.int64(32).tobytes()
    # default when the network definition carries no field (plugin.cpp:417)
    assert harness.harness_create_build(-1) == 0
    assert _text(harness.harness_serialised, raw, 16).startswith("parameters|") and bytes(raw[:8]) == np.int64(64).tobytes()
    # deserialisation: exactly the bytes the reference plugin would have written
    blob = (ctypes.c_ubyte * 8)(*np.int64(7).tobytes())
    assert harness.harness_create_runtime(b"parameters", FIELD_UNKNOWN, blob, 8) == 0
    assert _text(harness.harness_serialised, raw, 16) and bytes(raw[:8]) == np.int64(7).tobytes()
    # malformed serialised state is refused (nullptr), logged, and does not abort the process
    errors = harness.harness_logged_errors()
    assert harness.harness_create_runtime(b"parameters", FIELD_UNKNOWN, blob, 4) == 2
    assert harness.harness_create_runtime(b"params", FIELD_UNKNOWN, blob, 8) == 2
    assert harness.harness_create_runtime(b"parameters", FIELD_INT64, blob, 8) == 2
    zero = (ctypes.c_ubyte * 8)(*np.int64(0).tobytes())
    assert harness.harness_create_runtime(b"parameters", FIELD_UNKNOWN, zero, 8) == 2
    assert harness.harness_logged_errors() == errors + 4 and b"validation failed" in harness.harness_last_error()


@pytest.mark.parametrize("value_type,ok", [(TRT_FLOAT, True), (TRT_HALF, True), (TRT_BF16, True), (TRT_INT8, False), (TRT_INT32, False)])
def test_supports_format_combination(harness, value_type, ok):
    assert harness.harness_create_build(64) == 0
    arr = lambda xs: (ctypes.c_int * 6)(*xs)
    types = [value_type, TRT_INT64, TRT_INT64, value_type, value_type, value_type]
    for pos in range(6):
        expect = 1 if (ok or pos in (1, 2)) else 0
        assert harness.harness_supports(pos, arr(types), arr([LINEAR] * 6)) == expect, pos
    if ok:
        for pos in (0, 3, 4, 5):                                   # non-linear formats are refused
            fm = [LINEAR] * 6
            fm[pos] = CHW4
            assert harness.harness_supports(pos, arr(types), arr(fm)) == 0
        for pos in (3, 4, 5):                                      # all floating tensors share input 0's type
            mixed = list(types)
            mixed[pos] = TRT_HALF if value_type != TRT_HALF else TRT_FLOAT
            assert harness.harness_supports(pos, arr(mixed), arr([LINEAR] * 6)) == 0
        for pos in (1, 2):                                         # level tables are int64, nothing else
            bad = list(types)
            bad[pos] = TRT_INT32
            assert harness.harness_supports(pos, arr(bad), arr([LINEAR] * 6)) == 0
    assert harness.harness_supports(6, arr(types), arr([LINEAR] * 6)) == 0


def test_build_queries_and_malformed_network(harness):
    assert harness.harness_create_build(64) == 0
    vd = (ctypes.c_longlong * 4)(2, 18414, 8, 32)
    ld = (ctypes.c_longlong * 6)(2, 900, 8, 5, 4, 2)
    out = (ctypes.c_longlong * 7)()
    assert harness.harness_build_queries(vd, ld, TRT_BF16, out, 0) == 0
    assert list(out) == [0, 2, 900, 256, TRT_BF16, 0, 1]          # configure ok, [bs, Q, M*D], input 0's type, no workspace, 1 output
    errors = harness.harness_logged_errors()
    assert harness.harness_build_queries(vd, ld, TRT_HALF, out, 1) == 0
    assert out[0] == 1 and harness.harness_logged_errors() == errors + 1      # inconsistent head count: error return, no abort
    harness.harness_use_logger_finder()                            # setLoggerFinder route logs to the same place
    assert harness.harness_build_queries(vd, ld, TRT_HALF, out, 1) == 0 and harness.harness_logged_errors() == errors + 2


@pytest.mark.gpu
@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
def test_enqueue_equals_the_registered_op(harness, dt, cuda_device):
    import torch

    import codetr_b200 as cb
    from codetr_b200 import workloads as W

    tdt = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[dt]
    trt = {"f32": TRT_FLOAT, "f16": TRT_HALF, "bf16": TRT_BF16}[dt]
    keys = ("value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight")
    for name, batch in (("swinl_dec_1152x768", 2), ("r50_enc_608", 1)):
        inp = W.make_inputs(W.CONFIGS[name], batch=batch, out_of_range_frac=0.05)
        d = {}
        for k in keys:
            t = torch.from_numpy(getattr(inp, k))
            d[k] = t.to(cuda_device) if t.dtype == torch.int64 else t.to(device=cuda_device, dtype=tdt)
        want = torch.ops.codetr.multi_scale_deformable_attention(*(d[k] for k in keys), 64)
        got = torch.full_like(want, float("nan"))
        assert harness.harness_create_build(64) == 0
        stream = torch.cuda.Stream(device=cuda_device)
        stream.wait_stream(torch.cuda.current_stream(cuda_device))
        ptrs = (ctypes.c_void_p * 5)(*[d[k].data_ptr() for k in keys])
        vd = (ctypes.c_longlong * 4)(*d["value"].shape)
        ld = (ctypes.c_longlong * 6)(*d["sampling_loc"].shape)
        before = cb.launch_count()
        rc = harness.harness_enqueue(vd, ld, trt, ptrs, ctypes.c_void_p(got.data_ptr()), ctypes.c_void_p(stream.cuda_stream))
        stream.synchronize()
        assert rc == 0 and cb.launch_count() == before + 1
        assert torch.equal(got, want), name
    # a failing launch is reported by return code + log, not by abort (reference plugin.cpp:320-325 returns 1)
    errors = harness.harness_logged_errors()
    assert harness.harness_create_build(3) == 0                    # batch 2 is not divisible by min(2, 3)... (:924-926)
    rc = harness.harness_enqueue((ctypes.c_longlong * 4)(4, 10, 8, 32), (ctypes.c_longlong * 6)(4, 5, 8, 1, 4, 2), trt, ptrs,
                                 ctypes.c_void_p(got.data_ptr()), ctypes.c_void_p(stream.cuda_stream))
    assert rc == 1 and harness.harness_logged_errors() == errors + 1
