"""Parity metrics and gates (BASELINE.json north_star; SURVEY.md section 8(d) "Parity gates")."""
from __future__ import annotations

import numpy as np

FP32_REL_L2 = 1e-5       # fp32: ||out - ref||_2 / ||ref||_2 vs the fp32 reference on identical inputs
HALF_MAX_REL = 2e-3      # fp16: max|out - ref32| / max|ref32|, ref32 = fp32 reference on the upcast inputs
# bf16 keeps 8 significand bits: rounding the fp32 result once to bf16 already costs up to 2^-8 = 3.9e-3
# of the element (half an ulp at the bottom of a binade), so a 2e-3 gate cannot be met by ANY bf16 output --
# e.g. the reference's own seed-3 fixture rounds to 2.97e-3 with exact fp32 arithmetic (checked in
# tests/test_oracle.py::test_bf16_output_rounding_floor).  The bf16 gate is therefore "one output rounding":
# max-normalised error <= 2^-8 and every element within one bf16 ulp of the fp32 reference.
BF16_MAX_REL = 2.0 ** -8


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


def bf16_ulp_errors(out, ref) -> float:
    """Largest |out - ref| in units of the bf16 ulp of ref (ulp = 2^(floor(log2|ref|) - 7))."""
    out = np.asarray(out, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    mag = np.maximum(np.abs(ref), 2.0 ** -126)
    ulp = 2.0 ** (np.floor(np.log2(mag)) - 7)
    return float((np.abs(out - ref) / ulp).max())
