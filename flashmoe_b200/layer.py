"""Operator-style API on top of the drop-in surface: a forward-only `torch.nn.Module` that owns
the layer's parameters, uses them in place (no per-call re-upload like the reference's wrapper, python_bindings.cu:76-120),
supports the bias vectors and GELU the reference kernel has but cannot reach from Python, and can return the routing
decisions next to the output.  Inference only: no autograd (the reference is forward-only as well)."""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch

from .config import MoEConfig
from .runtime import MoEContext


class MoEOutput(NamedTuple):
    out: torch.Tensor                 # [B, T, H] bf16
    topk_idx: torch.Tensor            # [S, k] int32   expert picked per token, in pick order
    topk_weight: torch.Tensor         # [S, k] bf16    gateOut[t, e_j] (bf16-rounded softmax probability)
    slot: torch.Tensor                # [S, k] int32   slot in the (rank, expert) packet; >= capacity means dropped
    aux_loss: Optional[torch.Tensor] = None   # is_training = 1: f32 [2E+1] = gML[E] | gMeC[E] | load-balancing loss


class FlashMoELayer(torch.nn.Module):
    """One expert-parallel MoE layer: y[t] = sum_{e in topk(t), kept} (p~_e / sum_topk p) * FFN_e(x[t]).

    Parameters are held in the reference's layouts: `gate_weight` [H, E] (consumed flat as [E, H]),
    `expert_weight` [nLx, 2, P, H] (this rank's experts), optional `bias_up` [nLx, P] / `bias_down` [nLx, H].
    """

    def __init__(self, cfg: MoEConfig, rank: int = 0, world: int = 1, device: Optional[int] = None, group=None,
                 bias: bool = False):
        super().__init__()
        cfg.check_hot_path()
        self.cfg = cfg
        self.ctx = MoEContext(cfg, rank=rank, world=world, device=device, group=group)
        dev = self.ctx.device
        nlx = self.ctx.num_local_experts
        scale = cfg.H ** -0.5
        self.gate_weight = torch.nn.Parameter(torch.randn(cfg.H, cfg.E, device=dev).mul_(scale).bfloat16(),
                                              requires_grad=False)
        self.expert_weight = torch.nn.Parameter(torch.randn(nlx, 2, cfg.P, cfg.H, device=dev).mul_(scale).bfloat16(),
                                                requires_grad=False)
        if bias:
            self.bias_up = torch.nn.Parameter(torch.zeros(nlx, cfg.P, device=dev, dtype=torch.bfloat16), requires_grad=False)
            self.bias_down = torch.nn.Parameter(torch.zeros(nlx, cfg.H, device=dev, dtype=torch.bfloat16), requires_grad=False)
        else:
            self.bias_up = None
            self.bias_down = None

    @torch.no_grad()
    def forward(self, x: torch.Tensor, return_routing: bool = False):
        """x [B, T, H] bf16 on this layer's device with B*T == S.  Asynchronous on torch's current stream unless
        `return_routing` (reading the routing tables synchronises)."""
        out = self.ctx.forward(x, self.gate_weight, self.expert_weight, bias_up=self.bias_up, bias_down=self.bias_down)
        if not return_routing:
            return out
        self.ctx.synchronize()
        idx = torch.from_numpy(self.ctx.read("topk_idx"))
        w = torch.from_numpy(self.ctx.read("topk_w").view("int16")).view(torch.bfloat16)
        slot = torch.from_numpy(self.ctx.read("slot"))
        aux = torch.from_numpy(self.ctx.read("aux_loss")) if self.cfg.is_training else None
        return MoEOutput(out, idx, w, slot, aux)

    def extra_repr(self) -> str:
        c = self.cfg
        return (f"S={c.S}, H={c.H}, P={c.P}, E={c.E}, k={c.k}, act={'gelu' if c.hidden_act else 'relu'}, "
                f"rank={self.ctx.rank}/{self.ctx.world}")
