"""`flashmoe._C` surface (reference csrc/python_bindings.cu:194-217) over the C-ABI.

Same six functions, argument names and return keys:
    initialize(), finalize(), moe_forward(input, gate_weights, expert_weights) -> Tensor,
    get_compiled_config() -> {S,H,E,P,PX,Element_size}, get_bookkeeping() -> {nLx}, get_num_local_experts() -> int
Differences, all deliberate: errors raise RuntimeError instead of exit(1) (reference debug.cuh:19-43); one call =
one forward (the reference runs 32 warm-up + 32 timed launches inside moe_forward, python_bindings.cu:124 -- the
timing loop now lives in worker.py / bench.py); the caller's weights are used in place (no per-call re-upload).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from .runtime import MoEContext, env_rank_world

_lib.load()  # importing _C without the native library must fail loudly

_ctx: Optional[MoEContext] = None


def is_initialized() -> bool:
    return _ctx is not None


def context() -> MoEContext:
    if _ctx is None:
        raise RuntimeError("Must call initialize() before moe_forward")
    return _ctx


def initialize(config=None) -> None:
    """Create this process's context for the COMPILED config (or `config`, an extension for tests) and map peers.

    Rank / world size come from the launcher's environment (torchrun, or OMPI/PMI/SLURM like the reference's
    worker).  For world > 1 a torch.distributed process group is created if none exists (NCCL for CUDA tensors,
    gloo for the IPC-handle exchange)."""
    global _ctx
    if _ctx is not None:
        raise RuntimeError("initialize() called twice")  # reference asserts the same (bootstrap.cuh:537)
    rank, world, local = env_rank_world()
    ndev = torch.cuda.device_count()
    if ndev == 0:
        raise RuntimeError("no CUDA device visible: flashmoe_b200 has no CPU path")
    device = local % ndev
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist

        if not dist.is_initialized():
            dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world,
                                    device_id=torch.device("cuda", device))
    _ctx = MoEContext(config, rank=rank, world=world, device=device)


def finalize() -> None:
    global _ctx
    if _ctx is not None:
        _ctx.close()
        _ctx = None


def moe_forward(input: torch.Tensor, gate_weights: torch.Tensor, expert_weights: torch.Tensor) -> torch.Tensor:
    """MoE forward pass. Tensors must match compiled config dimensions (python_bindings.cu:17-65).
    Blocking, like the reference (cudaStreamSynchronize before return, :145)."""
    ctx = context()
    out = ctx.forward(input, gate_weights, expert_weights)
    ctx.synchronize()
    return out


def get_compiled_config() -> Dict[str, int]:
    if _ctx is not None:
        d = _ctx.dims
        return {"S": d["S"], "H": d["H"], "E": d["E"], "P": d["P"], "PX": d["PX"], "Element_size": d["element_size"]}
    return _lib.compiled_config().compiled_dict()


def get_bookkeeping() -> Dict[str, int]:
    return {"nLx": get_num_local_experts()}


def get_num_local_experts() -> int:
    return context().num_local_experts
