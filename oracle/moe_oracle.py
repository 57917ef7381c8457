"""ctypes front-end of oracle/moe_oracle.c + the multi-rank composition.

TEST INFRASTRUCTURE ONLY -- see the header of moe_oracle.c.  PARITY UNPINNED: the reference has no golden
vectors for this path and cannot be run here; the oracle restates the reference source.

Tensors cross this boundary as numpy arrays; bf16 is carried as uint16 bit patterns (`to_bits`/`from_bits`
convert from/to torch.bfloat16 without rounding).
"""
from __future__ import annotations

import ctypes
import dataclasses
import subprocess
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libmoe_oracle.so"
_lib: Optional[ctypes.CDLL] = None

_u16p = ctypes.POINTER(ctypes.c_uint16)
_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)


def build(force: bool = False) -> Path:
    """Compile the C restatement with the committed Makefile (gcc, no other dependencies)."""
    src = _HERE / "moe_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        L = ctypes.CDLL(str(_LIB_PATH))
        L.fmo_gate.restype = ctypes.c_int
        L.fmo_slots.restype = ctypes.c_int
        L.fmo_expert_ffn.restype = ctypes.c_int
        L.fmo_forward.restype = ctypes.c_int
        L.fmo_forward_sample.restype = ctypes.c_int
        L.fmo_forward_sample_w.restype = ctypes.c_int
        L.fmo_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray], ty):
    if a is None:
        return ctypes.cast(None, ty)
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ty)


# ----------------------------------------------------------------------------- bf16 helpers
def to_bits(t) -> np.ndarray:
    """torch.bfloat16 tensor (CPU) -> uint16 numpy array of the same shape (bit pattern, no rounding)."""
    import torch

    assert t.dtype == torch.bfloat16
    return t.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def from_bits(a: np.ndarray):
    import torch

    return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)


def bits_to_f32(a: np.ndarray) -> np.ndarray:
    return (a.astype(np.uint32) << 16).view(np.float32)


def f32_to_bits(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    out = np.empty(a.shape, dtype=np.uint16)
    lib().fmo_f32_to_bf16(_p(a, _f32p), _p(out, _u16p), ctypes.c_int64(a.size))
    return out


def num_threads() -> int:
    return int(lib().fmo_num_threads())


def set_num_threads(n: int) -> None:
    lib().fmo_set_num_threads(ctypes.c_int(n))


# ----------------------------------------------------------------------------- operand reinterpretation (A.1)
def gate_weights_effective(gate_weights_bits: np.ndarray, E: int, H: int) -> np.ndarray:
    """[H,E] contiguous tensor viewed flat as [E,H] (python_bindings.cu:93-99, moe.cuh:107-109)."""
    assert gate_weights_bits.shape == (H, E)
    return np.ascontiguousarray(gate_weights_bits).reshape(-1).reshape(E, H)


def split_expert_weights(expert_weights_bits: np.ndarray):
    """[n,2,P,H] -> (w_up [n,P,H], w_down_eff [n,H,P]); the down block is REINTERPRETED, not transposed
    (python_bindings.cu:104-119, moe.cuh:111-116)."""
    n, two, P, H = expert_weights_bits.shape
    assert two == 2
    w_up = np.ascontiguousarray(expert_weights_bits[:, 0])
    w_down_eff = np.ascontiguousarray(expert_weights_bits[:, 1]).reshape(n, -1).reshape(n, H, P)
    return w_up, w_down_eff


# ----------------------------------------------------------------------------- results
@dataclasses.dataclass
class OracleResult:
    out: np.ndarray        # bf16 bits [S,H]
    topk_idx: np.ndarray   # int32 [S,k]
    slot: np.ndarray       # int32 [S,k]  position among all selections of that expert on this rank
    kept: np.ndarray       # int32 [S,k]  1 iff slot < EC
    counts: np.ndarray     # int32 [E]    all selections (incl. dropped)
    mcw: np.ndarray        # float32 [S]
    gate_out: np.ndarray   # bf16 bits [S,E]
    logits: np.ndarray     # float32 [S,E]
    ambiguous: np.ndarray  # bool [S]  tokens whose top-k indices are not robust to fp32 summation order

    def out_f32(self) -> np.ndarray:
        return bits_to_f32(self.out)


def ambiguity_flags(logits: np.ndarray, abs_sum: np.ndarray, k: int) -> np.ndarray:
    """Tokens for which a different (still valid) fp32 accumulation order could change the ordered top-k list.

    A token is flagged when two logits among the sorted positions 1..k+1 are closer than 16 ulp of the
    largest |x.w| absolute sum (2^-20 * abs_sum), or when any expert sits within 1e-3 of the exp underflow
    threshold (the ftz flush at 2^-126 decides whether its probability is exactly 0, Appendix A.4).
    """
    S, E = logits.shape
    srt = -np.sort(-logits.astype(np.float64), axis=1)
    top = srt[:, : min(k + 1, E)]
    gaps = top[:, :-1] - top[:, 1:] if top.shape[1] > 1 else np.full((S, 1), np.inf)
    tol = abs_sum.astype(np.float64) * 2.0 ** -20
    near_tie = (gaps < tol[:, None]).any(axis=1)
    t = (logits.astype(np.float64) - srt[:, :1]) * 1.4426950408889634
    near_flush = (np.abs(t + 126.0) < 1e-3).any(axis=1)
    return near_tie | near_flush


def forward(x: np.ndarray, wg_eff: np.ndarray, w_up: np.ndarray, w_down_eff: np.ndarray, *, k: int, EC: int,
            act: int = 0, b_up: Optional[np.ndarray] = None, b_down: Optional[np.ndarray] = None) -> OracleResult:
    """One rank's tokens through the whole layer with ALL experts' weights given (bf16 bit arrays)."""
    S, H = x.shape
    E, P, H2 = w_up.shape
    assert H2 == H and wg_eff.shape == (E, H) and w_down_eff.shape == (E, H, P)
    for a in (x, wg_eff, w_up, w_down_eff):
        assert a.dtype == np.uint16
    x = np.ascontiguousarray(x)
    out = np.zeros((S, H), dtype=np.uint16)
    topk = np.zeros((S, k), dtype=np.int32)
    slot = np.zeros((S, k), dtype=np.int32)
    kept = np.zeros((S, k), dtype=np.int32)
    counts = np.zeros((E,), dtype=np.int32)
    mcw = np.zeros((S,), dtype=np.float32)
    gate_out = np.zeros((S, E), dtype=np.uint16)
    logits = np.zeros((S, E), dtype=np.float32)
    abs_sum = np.zeros((S,), dtype=np.float32)
    rc = lib().fmo_forward(
        _p(x, _u16p), _p(np.ascontiguousarray(wg_eff), _u16p), _p(np.ascontiguousarray(w_up), _u16p),
        _p(np.ascontiguousarray(w_down_eff), _u16p),
        _p(None if b_up is None else np.ascontiguousarray(b_up), _u16p),
        _p(None if b_down is None else np.ascontiguousarray(b_down), _u16p),
        ctypes.c_int(S), ctypes.c_int(H), ctypes.c_int(P), ctypes.c_int(E), ctypes.c_int(k), ctypes.c_int(EC),
        ctypes.c_int(act), _p(out, _u16p), _p(topk, _i32p), _p(slot, _i32p), _p(kept, _i32p), _p(counts, _i32p),
        _p(mcw, _f32p), _p(gate_out, _u16p), _p(logits, _f32p), _p(abs_sum, _f32p))
    if rc != 0:
        raise RuntimeError(f"fmo_forward failed with code {rc}")
    return OracleResult(out, topk, slot, kept, counts, mcw, gate_out, logits, ambiguity_flags(logits, abs_sum, k))


def forward_sample(x: np.ndarray, wg_eff: np.ndarray, w_up: np.ndarray, w_down_eff: np.ndarray, sample: np.ndarray, *,
                   k: int, EC: int, act: int = 0, b_up: Optional[np.ndarray] = None,
                   b_down: Optional[np.ndarray] = None, topk_w_given: Optional[np.ndarray] = None,
                   mcw_given: Optional[np.ndarray] = None) -> OracleResult:
    """Like `forward`, but the expert FFN + combine run only for the tokens listed in `sample` (routing, slots and
    capacity drops are still computed over all S tokens).  `out` of the result is [len(sample), H] in sample order;
    every other field covers all S tokens.  Keeps full-size shapes (d_model 4096, ffn 14336) at seconds of CPU.

    `topk_w_given` (bf16 bits [len(sample), k]) / `mcw_given` (float32 [len(sample)]): combine with these router weights
    instead of the oracle's own (see fmo_forward_sample_w: separates 1-ulp differences of the bf16 gate probabilities,
    which depend on the router GEMM's summation order, from the FFN + combine arithmetic)."""
    S, H = x.shape
    E, P, H2 = w_up.shape
    assert H2 == H and wg_eff.shape == (E, H) and w_down_eff.shape == (E, H, P)
    sample = np.ascontiguousarray(sample, dtype=np.int32)
    n = int(sample.size)
    out = np.zeros((n, H), dtype=np.uint16)
    topk = np.zeros((S, k), dtype=np.int32)
    slot = np.zeros((S, k), dtype=np.int32)
    kept = np.zeros((S, k), dtype=np.int32)
    counts = np.zeros((E,), dtype=np.int32)
    mcw = np.zeros((S,), dtype=np.float32)
    gate_out = np.zeros((S, E), dtype=np.uint16)
    logits = np.zeros((S, E), dtype=np.float32)
    abs_sum = np.zeros((S,), dtype=np.float32)
    wv = None if topk_w_given is None else np.ascontiguousarray(topk_w_given, dtype=np.uint16).reshape(n, k)
    mv = None if mcw_given is None else np.ascontiguousarray(mcw_given, dtype=np.float32).reshape(n)
    rc = lib().fmo_forward_sample_w(
        _p(np.ascontiguousarray(x), _u16p), _p(np.ascontiguousarray(wg_eff), _u16p),
        _p(np.ascontiguousarray(w_up), _u16p), _p(np.ascontiguousarray(w_down_eff), _u16p),
        _p(None if b_up is None else np.ascontiguousarray(b_up), _u16p),
        _p(None if b_down is None else np.ascontiguousarray(b_down), _u16p),
        ctypes.c_int(S), ctypes.c_int(H), ctypes.c_int(P), ctypes.c_int(E), ctypes.c_int(k), ctypes.c_int(EC),
        ctypes.c_int(act), _p(sample, _i32p), ctypes.c_int(n), _p(wv, _u16p), _p(mv, _f32p), _p(out, _u16p),
        _p(topk, _i32p), _p(slot, _i32p),
        _p(kept, _i32p), _p(counts, _i32p), _p(mcw, _f32p), _p(gate_out, _u16p), _p(logits, _f32p), _p(abs_sum, _f32p))
    if rc != 0:
        raise RuntimeError(f"fmo_forward_sample_w failed with code {rc}")
    return OracleResult(out, topk, slot, kept, counts, mcw, gate_out, logits, ambiguity_flags(logits, abs_sum, k))


def router_weight_ulps(dev_topk_w: np.ndarray, ref: OracleResult, sample: np.ndarray) -> np.ndarray:
    """Distance in bf16 ulps between the device's top-k router weights (bf16 bits [len(sample), k]) and the oracle's
    gateOut[t, e_j] for the sampled tokens (meaningful where the top-k indices agree; probabilities are positive, so the
    ulp distance is the difference of the bit patterns)."""
    t = np.asarray(sample, dtype=np.int64)
    ref_w = ref.gate_out[t[:, None], ref.topk_idx[t]]
    return np.abs(np.asarray(dev_topk_w, dtype=np.uint16).astype(np.int64) - ref_w.astype(np.int64))


def aux_loss(x: np.ndarray, wg_eff: np.ndarray, *, k: int, EC: int):
    """Training-mode auxiliary loss of one rank's tokens: (gML [E], gMeC [E], loss) as float32
    (reference moe/gate.cuh:608-635,698-706,763-773; see fmo_aux_loss)."""
    S, H = x.shape
    E = wg_eff.shape[0]
    logits = np.zeros((S, E), dtype=np.float32)
    probs = np.zeros((S, E), dtype=np.float32)
    gate_out = np.zeros((S, E), dtype=np.uint16)
    topk = np.zeros((S, k), dtype=np.int32)
    mcw = np.zeros((S,), dtype=np.float32)
    abs_sum = np.zeros((S,), dtype=np.float32)
    rc = lib().fmo_gate(_p(np.ascontiguousarray(x), _u16p), _p(np.ascontiguousarray(wg_eff), _u16p), ctypes.c_int(S),
                        ctypes.c_int(H), ctypes.c_int(E), ctypes.c_int(k), _p(logits, _f32p), _p(probs, _f32p),
                        _p(gate_out, _u16p), _p(topk, _i32p), _p(mcw, _f32p), _p(abs_sum, _f32p))
    if rc != 0:
        raise RuntimeError(f"fmo_gate failed with code {rc}")
    slot = np.zeros((S, k), dtype=np.int32)
    kept = np.zeros((S, k), dtype=np.int32)
    counts = np.zeros((E,), dtype=np.int32)
    rc = lib().fmo_slots(_p(topk, _i32p), ctypes.c_int(S), ctypes.c_int(E), ctypes.c_int(k), ctypes.c_int(EC),
                         _p(slot, _i32p), _p(kept, _i32p), _p(counts, _i32p))
    if rc != 0:
        raise RuntimeError(f"fmo_slots failed with code {rc}")
    gml = np.zeros((E,), dtype=np.float32)
    gmec = np.zeros((E,), dtype=np.float32)
    loss = np.zeros((1,), dtype=np.float32)
    lib().fmo_aux_loss(_p(probs, _f32p), _p(counts, _i32p), ctypes.c_int(S), ctypes.c_int(E), _p(gml, _f32p),
                       _p(gmec, _f32p), _p(loss, _f32p))
    return gml, gmec, float(loss[0])


def expert_ffn(rows: np.ndarray, w_up: np.ndarray, w_down_eff: np.ndarray, act: int = 0,
               b_up: Optional[np.ndarray] = None, b_down: Optional[np.ndarray] = None):
    """h, y (bf16 bits) for a packet of rows through one expert (Appendix A.6)."""
    R, H = rows.shape
    P = w_up.shape[0]
    h = np.zeros((R, P), dtype=np.uint16)
    y = np.zeros((R, H), dtype=np.uint16)
    rc = lib().fmo_expert_ffn(_p(np.ascontiguousarray(rows), _u16p), ctypes.c_int64(R), ctypes.c_int(H),
                              ctypes.c_int(P), _p(np.ascontiguousarray(w_up), _u16p),
                              _p(np.ascontiguousarray(w_down_eff), _u16p),
                              _p(None if b_up is None else np.ascontiguousarray(b_up), _u16p),
                              _p(None if b_down is None else np.ascontiguousarray(b_down), _u16p),
                              ctypes.c_int(act), _p(h, _u16p), _p(y, _u16p))
    if rc != 0:
        raise RuntimeError(f"fmo_expert_ffn failed with code {rc}")
    return h, y


def forward_world(xs: Sequence[np.ndarray], gate_weights: Sequence[np.ndarray],
                  expert_weights: Sequence[np.ndarray], *, k: int, EC: int, act: int = 0) -> List[OracleResult]:
    """Expert-parallel world of W ranks (Appendix A.8).

    xs[r] [S,H]; gate_weights[r] [H,E] (rank r gates its own tokens with its own tensor, as the kernel does);
    expert_weights[r] [nLx,2,P,H] = rank r's local experts; expert e lives on rank e // nLx at index e % nLx
    (bootstrap.cuh:35-52).  Capacity EC applies per (source rank, expert) packet, so each rank's tokens see an
    independent single-rank problem over the concatenated expert weights.
    """
    W = len(xs)
    ups, downs = zip(*(split_expert_weights(w) for w in expert_weights))
    w_up = np.concatenate(ups, axis=0)
    w_down_eff = np.concatenate(downs, axis=0)
    E = w_up.shape[0]
    res = []
    for r in range(W):
        H = xs[r].shape[1]
        res.append(forward(xs[r], gate_weights_effective(gate_weights[r], E, H), w_up, w_down_eff, k=k, EC=EC,
                           act=act))
    return res
