"""Parity metrics and gates (BASELINE.json north_star; SURVEY.md section 8(d) "Parity gates")."""
from __future__ import annotations

import numpy as np

FP32_REL_L2 = 1e-5       # fp32: ||out - ref||_2 / ||ref||_2 vs the fp32 reference on identical inputs
HALF_MAX_REL = 2e-3      # fp16/bf16: max|out - ref32| / max|ref32|, ref32 = fp32 reference on the upcast inputs


def rel_l2(out, ref) -> float:
    out = np.asarray(out, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.linalg.norm(ref.ravel())
    return float(np.linalg.norm((out - ref).ravel()) / (den if den > 0 else 1.0))


def max_rel(out, ref) -> float:
    """The reference's own normalisation (tests/test_multi_scale_deformable_attention.py:485):
    max abs error over the max abs reference value."""
    out = np.asarray(out, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.abs(ref).max() if ref.size else 0.0
    return float(np.abs(out - ref).max() / (den if den > 0 else 1.0)) if ref.size else 0.0


def max_abs(out, ref) -> float:
    out = np.asarray(out, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(out - ref).max()) if ref.size else 0.0
