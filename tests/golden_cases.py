"""Seeded input definitions shared by tests/golden/make_golden.py (which runs the reference on
them) and the tests (which run the oracle and the CUDA path on them)."""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import Callable, Dict, List

import numpy as np

import codetr_b200
from codetr_b200 import workloads as W

ARRAY_KEYS = ("value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight")


def inputs_digest(arrs: Dict[str, np.ndarray]) -> str:
    h = hashlib.sha256()
    for k in ARRAY_KEYS:
        a = np.ascontiguousarray(arrs[k])
        h.update(k.encode())
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


@dataclass
class Case:
    name: str
    build: Callable[[], Dict[str, np.ndarray]]
    store_inputs: bool = True
    note: str = ""


def _pack(value, shapes, loc, weight) -> Dict[str, np.ndarray]:
    shapes = np.asarray(shapes, dtype=np.int64).reshape(-1, 2)
    starts = np.concatenate([[0], np.cumsum(shapes[:, 0] * shapes[:, 1])[:-1]]).astype(np.int64)
    return dict(value=np.ascontiguousarray(value, dtype=np.float32), spatial_shapes=shapes, level_start_index=starts,
                sampling_loc=np.ascontiguousarray(loc, dtype=np.float32),
                attn_weight=np.ascontiguousarray(weight, dtype=np.float32))


def _from_inputs(inp: W.Inputs) -> Dict[str, np.ndarray]:
    return dict(value=inp.value, spatial_shapes=inp.spatial_shapes, level_start_index=inp.level_start_index,
                sampling_loc=inp.sampling_loc, attn_weight=inp.attn_weight)


def _ref_seed3() -> Dict[str, np.ndarray]:
    """The reference's known-shape fixture (tests/test_multi_scale_deformable_attention.py:286-299):
    N=1, M=2, D=2, Lq=2, L=2 [(6,4),(3,2)], P=2, torch.manual_seed(3), value=rand*0.01, weights
    normalised over L*P.  torch's CPU RNG stream is used on purpose (the arrays are stored)."""
    import torch

    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    shapes = [(6, 4), (3, 2)]
    S = sum(h * w for h, w in shapes)
    torch.manual_seed(3)
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2)
    w = torch.rand(N, Lq, M, L, P) + 1e-5
    w /= w.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return _pack(value.numpy(), shapes, loc.numpy(), w.numpy())


def _ref_forward_small() -> Dict[str, np.ndarray]:
    """Shape of the reference's random forward test (tests:14-62): B=2, M=4, Q=8, D=16,
    L=3 (32x32, 16x16, 8x8), P=4, uniform inputs."""
    rng = np.random.default_rng(101)
    shapes = [(32, 32), (16, 16), (8, 8)]
    S = sum(h * w for h, w in shapes)
    return _pack(rng.random((2, S, 4, 16), dtype=np.float32), shapes, rng.random((2, 8, 4, 3, 4, 2), dtype=np.float32),
                 rng.random((2, 8, 4, 3, 4), dtype=np.float32))


def _ref_mid() -> Dict[str, np.ndarray]:
    """BASELINE.json configs[0]: the reference's test_benchmark_performance shape (tests:417-428):
    N=1, M=8, D=64, Lq=100, L=4 [(64,64),(32,32),(16,16),(8,8)], P=4, uniform inputs, weights
    normalised over L*P."""
    wl = W.CONFIGS["ref_test_mid_fp32"]
    rng = np.random.default_rng(wl.seed)
    S = wl.S
    value = rng.random((1, S, 8, 64), dtype=np.float32)
    loc = rng.random((1, 100, 8, 4, 4, 2), dtype=np.float32)
    w = rng.random((1, 100, 8, 4, 4), dtype=np.float32)
    w /= w.sum(-1, keepdims=True).sum(-2, keepdims=True)
    return _pack(value, wl.shapes, loc, w)


def _edge_borders() -> Dict[str, np.ndarray]:
    """Hand-placed locations on odd, non-square levels: exactly on 0 and 1, on pixel centres, half a
    pixel outside (x_pix = -1 and x_pix = W exactly, the open ends of the range test), far outside,
    and negative."""
    rng = np.random.default_rng(202)
    shapes = [(19, 13), (10, 7), (5, 4), (1, 3)]
    L, M, D, P = len(shapes), 2, 32, 4
    S = sum(h * w for h, w in shapes)
    value = rng.standard_normal((1, S, M, D), dtype=np.float32)
    special = []
    for h, w in shapes:
        xs = [0.0, 1.0, 0.5 / w, 1.0 - 0.5 / w, -0.5 / w, (w + 0.5) / w, 1.5 / w, -0.25 / w, (w + 0.25) / w, 0.5, -3.0, 4.0]
        ys = [0.0, 1.0, 0.5 / h, 1.0 - 0.5 / h, -0.5 / h, (h + 0.5) / h, 1.5 / h, -0.25 / h, (h + 0.25) / h, 0.5, -3.0, 4.0]
        special.append((xs, ys))
    n = len(special[0][0])
    Q = n * n // P + 1
    loc = np.zeros((1, Q, M, L, P, 2), dtype=np.float32)
    for l in range(L):
        xs, ys = special[l]
        pairs = [(x, y) for x in xs for y in ys]
        for i in range(Q * P):
            x, y = pairs[i % len(pairs)]
            loc[0, i // P, 0, l, i % P] = (x, y)
            x2, y2 = pairs[(i * 7 + 3) % len(pairs)]
            loc[0, i // P, 1, l, i % P] = (x2, y2)
    w = rng.random((1, Q, M, L, P), dtype=np.float32) + 0.05
    return _pack(value, shapes, loc, w)


def _tiny_pyramid(kind: str, q: int, seed: int, oor: float, batch: int = 2) -> Dict[str, np.ndarray]:
    wl = W.Workload(name="tiny", shapes=tuple(W.pyramid_shapes(64, 96)), num_queries=q, batch=batch, kind=kind, seed=seed)
    return _from_inputs(W.make_inputs(wl, out_of_range_frac=oor))


def _odd_dims() -> Dict[str, np.ndarray]:
    """Shapes the vector kernels do not cover (D=5, M=3, P=3, L=2): exercises the generic kernel."""
    rng = np.random.default_rng(303)
    shapes = [(7, 9), (4, 5)]
    S = sum(h * w for h, w in shapes)
    loc = rng.uniform(-0.1, 1.1, size=(2, 11, 3, 2, 3, 2)).astype(np.float32)
    return _pack(rng.standard_normal((2, S, 3, 5), dtype=np.float32), shapes, loc, rng.random((2, 11, 3, 2, 3), dtype=np.float32))


def cases() -> List[Case]:
    return [
        Case("ref_seed3", _ref_seed3, note="reference tests:286-299 fixture"),
        Case("ref_forward_small", _ref_forward_small, note="reference tests:14-62 shape"),
        Case("ref_mid", _ref_mid, store_inputs=False, note="BASELINE configs[0] / reference tests:417-428 shape"),
        Case("edge_borders", _edge_borders, note="border / out-of-range semantics on odd levels"),
        Case("codino_enc_tiny", lambda: _tiny_pyramid("encoder", 0, 11, 0.05, batch=1), note="Co-DINO encoder, 96x64 image"),
        Case("codino_dec_tiny", lambda: _tiny_pyramid("decoder", 50, 12, 0.0), note="Co-DINO decoder, 50 queries"),
        Case("odd_dims", _odd_dims, note="D=5, M=3, P=3 (generic kernel)"),
    ]


def case_by_name(name: str) -> Case:
    for c in cases():
        if c.name == name:
            return c
    raise KeyError(name)
