"""Plain torch CPU MoE forward: the "non-optimised baseline" BASELINE.json's north_star asks to be timed beside
the GPU path, and an independent cross-check of oracle/moe_oracle.c.

TEST INFRASTRUCTURE ONLY (see oracle/moe_oracle.c).  PARITY UNPINNED for the same reason.

Same semantics as SURVEY.md Appendix A (fp32 matmuls on bf16 operands, RNE at the h / y / gateOut store points,
strict-'>' top-k on fp32 probabilities, ascending-token capacity slots, bf16 weighted combine) but written the way
a torch user would write it: `torch.softmax`, `torch.matmul` (MKL blocked summation order), argmax rounds.  It
therefore differs from the C oracle only by fp32 summation order / libm-vs-MKL exp, which is exactly the
freedom a second implementation of the reference has.
"""
from __future__ import annotations

import time
from typing import Dict, Optional, Tuple

import torch


def _rne(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def _act(t: torch.Tensor, act: int) -> torch.Tensor:
    if act == 0:
        return torch.relu(t)
    if act == 1:
        return torch.nn.functional.gelu(t)  # erf form, like cutlass::epilogue::thread::GELU
    return t


def route(x: torch.Tensor, gate_weights: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """x [S,H] bf16, gate_weights [H,E] bf16 -> (probs f32 [S,E], topk_idx i64 [S,k], mcw f32 [S])."""
    S, H = x.shape
    E = gate_weights.shape[1]
    wg_eff = gate_weights.contiguous().view(-1).view(E, H)  # reinterpretation (python_bindings.cu:93-99)
    logits = x.float() @ wg_eff.float().t()
    probs = torch.softmax(logits, dim=-1)
    work = probs.clone()
    picks, mcw = [], torch.zeros(S)
    for _ in range(k):  # k rounds of argmax; torch.argmax returns the first maximal index == strict '>' scan
        idx = torch.argmax(work, dim=-1)
        val = work.gather(1, idx[:, None])[:, 0]
        picks.append(idx)
        mcw = mcw + val
        work.scatter_(1, idx[:, None], float("-inf"))
    return probs, torch.stack(picks, dim=1), mcw


def moe_forward_cpu(x: torch.Tensor, gate_weights: torch.Tensor, expert_weights: torch.Tensor, *, k: int, EC: int,
                    act: int = 0, bias_up: Optional[torch.Tensor] = None,
                    bias_down: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Single-rank (all experts local) forward.  x [S,H], gate_weights [H,E], expert_weights [E,2,P,H], bf16 CPU."""
    S, H = x.shape
    E, _, P, _ = expert_weights.shape
    probs, topk, mcw = route(x, gate_weights, k)
    p_tilde = _rne(probs)  # gateOut is stored in bf16 (gate.cuh:590-605)
    # capacity slots in ascending token order (Appendix A.5)
    flat_e = topk.reshape(-1)
    # position of each (token, j) among all selections of its expert, token-major order
    order_onehot = torch.nn.functional.one_hot(flat_e, E)
    excl = torch.cumsum(order_onehot, dim=0) - order_onehot
    slot = excl.gather(1, flat_e[:, None])[:, 0].view(S, k)
    kept = slot < EC
    xf = x.float()
    out_terms = torch.zeros(k, S, H)
    contributes = torch.zeros(k, S, dtype=torch.bool)
    for e in range(E):
        w_up = expert_weights[e, 0].float()                                  # [P,H]
        w_down_eff = expert_weights[e, 1].contiguous().view(-1).view(H, P).float()  # reinterpretation
        sel = (topk == e) & kept                                             # [S,k]
        tok, j = sel.nonzero(as_tuple=True)
        if tok.numel() == 0:
            continue
        h = xf[tok] @ w_up.t()
        if bias_up is not None:
            h = h + bias_up[e].float()
        h = _rne(_act(h, act))
        y = h @ w_down_eff.t()
        if bias_down is not None:
            y = y + bias_down[e].float()
        y = _rne(y)
        if k == 1:
            out_terms[0, tok] = y
        else:
            q = _rne(y / mcw[tok, None])
            out_terms[j, tok] = _rne(p_tilde[tok, e][:, None] * q)
        contributes[j, tok] = True
    if k == 1:
        out = out_terms[0]
    else:
        out = torch.zeros(S, H)
        for j in range(k):  # bf16 accumulation, ascending j
            out = torch.where(contributes[j][:, None], _rne(out + out_terms[j]), out)
    return {"out": out.to(torch.bfloat16), "topk_idx": topk.to(torch.int32), "slot": slot.to(torch.int32),
            "kept": kept, "mcw": mcw, "probs": probs}


def time_cpu_forward(x, gate_weights, expert_weights, *, k: int, EC: int, act: int = 0, iters: int = 3,
                     warmup: int = 1) -> Tuple[float, int]:
    """Wall-clock seconds per forward (mean of `iters` after `warmup`) and the thread count used."""
    threads = torch.get_num_threads()
    for _ in range(warmup):
        moe_forward_cpu(x, gate_weights, expert_weights, k=k, EC=EC, act=act)
    t0 = time.perf_counter()
    for _ in range(iters):
        moe_forward_cpu(x, gate_weights, expert_weights, k=k, EC=EC, act=act)
    return (time.perf_counter() - t0) / iters, threads
