"""Shared helpers for the parity tests: seeded synthetic inputs, the oracle call, and the acceptance rule.

Acceptance rule (BASELINE.json north_star "bit-exact top-k expert indices, combined activations
within 1e-3 relative bf16 tolerance"), written out once here and used by every parity test:
  * top-k indices: bit-exact, in pick order, for every token the oracle does not flag as ambiguous (two candidate
    logits closer than 16 ulp of the largest |x.w| partial-sum magnitude, where any fp32 summation order is a valid
    implementation of the reference); flagged tokens must stay below 0.1 % of the tokens.
  * slots / counts: exact integer equality whenever the routing of all earlier tokens agrees.
  * outputs (bf16): relative Frobenius error <= 1e-3; no NaN/Inf; per element |d| <= 1e-2*|ref| + 1e-3*max|ref| for
    >= 99.9 % of the elements (bf16 carries 8 significant bits: two correct fp32 summation orders may round an
    element to adjacent bf16 values, a 2^-8 relative step, so a per-element 1e-3 bound is not a property even the
    reference has against itself).
"""
from __future__ import annotations

import numpy as np
import torch

from flashmoe_b200.config import MoEConfig
from oracle import moe_oracle as mo

REL_FROBENIUS_TOL = 1e-3
ELEM_RTOL, ELEM_ATOL_FRAC, ELEM_OK_FRACTION = 1e-2, 1e-3, 0.999
MAX_AMBIGUOUS_FRACTION = 1e-3


def make_inputs(cfg: MoEConfig, seed: int = 0, scaled: bool = True, n_local: int | None = None, bias: bool = False):
    """(x [mb,seq,H], gate_weights [H,E], expert_weights [n,2,P,H], bias_up, bias_down) as CPU bf16 tensors.
    `scaled` multiplies the weights by H^-1/2 (non-degenerate softmax, O(1) activations -- SURVEY.md 8d)."""
    g = torch.Generator().manual_seed(seed)
    n = cfg.E if n_local is None else n_local
    x = torch.randn(cfg.mini_batch, cfg.sequence_len, cfg.H, generator=g)
    wg = torch.randn(cfg.H, cfg.E, generator=g)
    we = torch.randn(n, 2, cfg.P, cfg.H, generator=g)
    if scaled:
        wg, we = wg * cfg.H ** -0.5, we * cfg.H ** -0.5
    bu = bd = None
    if bias:
        bu = (torch.randn(n, cfg.P, generator=g) * 0.1).bfloat16()
        bd = (torch.randn(n, cfg.H, generator=g) * 0.1).bfloat16()
    return x.bfloat16(), wg.bfloat16(), we.bfloat16(), bu, bd


def run_oracle(cfg: MoEConfig, x, wg, we, bu=None, bd=None) -> mo.OracleResult:
    up, down = mo.split_expert_weights(mo.to_bits(we))
    return mo.forward(mo.to_bits(x.reshape(cfg.S, cfg.H)), mo.gate_weights_effective(mo.to_bits(wg), cfg.E, cfg.H), up,
                      down, k=cfg.k, EC=cfg.EC, act=cfg.hidden_act,
                      b_up=None if bu is None else mo.to_bits(bu), b_down=None if bd is None else mo.to_bits(bd))


def check_topk(got_idx: np.ndarray, ref: mo.OracleResult):
    S = got_idx.shape[0]
    mism = (got_idx != ref.topk_idx).any(axis=1)
    hard = mism & ~ref.ambiguous
    assert ref.ambiguous.mean() <= MAX_AMBIGUOUS_FRACTION or S < 2000, f"{int(ref.ambiguous.sum())} ambiguous tokens of {S}"
    assert not hard.any(), (f"top-k indices differ on {int(hard.sum())} unambiguous tokens, e.g. token "
                            f"{int(np.flatnonzero(hard)[0])}: got {got_idx[np.flatnonzero(hard)[0]]} "
                            f"ref {ref.topk_idx[np.flatnonzero(hard)[0]]}")
    return mism


def check_output(got_bits: np.ndarray, ref_bits: np.ndarray, rows_ok: np.ndarray | None = None, what: str = "out"):
    g = mo.bits_to_f32(got_bits).astype(np.float64)
    r = mo.bits_to_f32(ref_bits).astype(np.float64)
    if rows_ok is not None:
        g, r = g[rows_ok], r[rows_ok]
    assert np.isfinite(g).all(), f"{what}: non-finite values"
    denom = max(np.linalg.norm(r), 1e-30)
    relf = np.linalg.norm(g - r) / denom
    assert relf <= REL_FROBENIUS_TOL, f"{what}: relative Frobenius error {relf:.3e} > {REL_FROBENIUS_TOL}"
    tol = ELEM_RTOL * np.abs(r) + ELEM_ATOL_FRAC * max(np.abs(r).max(), 1e-30)
    ok = (np.abs(g - r) <= tol).mean()
    assert ok >= ELEM_OK_FRACTION, f"{what}: only {ok:.5f} of the elements within tolerance"
    return relf
