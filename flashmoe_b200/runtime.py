"""Host-side runtime above the C-ABI: one `MoEContext` per process / GPU.

Mirrors what the reference does in `flashmoe::initialize()` + `moe_forward` (csrc/include/flashmoe/bootstrap.cuh:278-547,
csrc/python_bindings.cu:17-151) with torch used only for device memory, streams and the out-of-band exchange of
CUDA IPC handles (`torch.distributed`, any backend) -- plumbing, not the product.  All math runs in the native
library; nothing here falls back to torch ops.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .config import MoEConfig

_BUFFER_SPECS = {
    "topk_idx": (_lib.BUF_TOPK_IDX, np.int32),
    "topk_w": (_lib.BUF_TOPK_W, np.uint16),
    "mcw": (_lib.BUF_MCW, np.float32),
    "slot": (_lib.BUF_SLOT, np.int32),
    "counts": (_lib.BUF_COUNTS, np.int32),
    "recv_x": (_lib.BUF_RECV_X, np.uint16),
    "hidden": (_lib.BUF_HIDDEN, np.uint16),
    "ret_y": (_lib.BUF_RET_Y, np.uint16),
    "gate_out": (_lib.BUF_GATE_OUT, np.uint16),
    "recv_cnt": (_lib.BUF_RECV_CNT, np.int32),
    "trace": (_lib.BUF_TRACE, np.uint64),
    "aux_loss": (_lib.BUF_AUX_LOSS, np.float32),
}


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's env, falling back to the launchers the reference's worker reads
    (OMPI / PMI / SLURM, reference flashmoe/worker.py:24-29)."""
    def first(*names: str, default: str) -> int:
        for n in names:
            if n in os.environ:
                return int(os.environ[n])
        return int(default)

    rank = first("RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID", default="0")
    world = first("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS", default="1")
    local = first("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID", default=str(rank))
    return rank, world, local


def exchange_blobs(mine: bytes, world: int, group=None) -> bytes:
    """All-gather one fixed-size byte string per rank over torch.distributed (any backend) and return them
    concatenated in rank order -- the out-of-band channel for the CUDA IPC handles of the symmetric slabs."""
    import torch.distributed as dist

    if not dist.is_initialized():
        raise RuntimeError("world > 1 needs an initialised torch.distributed process group")
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    if any(len(g) != len(mine) for g in gathered):
        raise RuntimeError("peers exported handles of different sizes")
    return b"".join(gathered)


def _check_tensor(name: str, t: torch.Tensor, device: torch.device) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be CUDA tensor")
    if t.device != device:
        raise RuntimeError(f"{name} is on {t.device}, context is on {device}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != torch.bfloat16:
        raise RuntimeError(f"{name} must be torch.bfloat16 (compiled Element), got {t.dtype}")
    if t.data_ptr() % 16:
        raise RuntimeError(f"{name} must start on a 16-byte boundary (TMA tensor maps, bulk row copies and the vector "
                           f"accesses need it); got storage offset {t.storage_offset()}")


class MoEContext:
    """Owns the native context (workspaces, symmetric slab, peer mappings) for one rank."""

    def __init__(self, cfg: Optional[MoEConfig] = None, rank: int = 0, world: int = 1,
                 device: Optional[int] = None, group=None, timeout_ms: Optional[int] = None,
                 symmetric: Optional[str] = None):
        """`symmetric` selects how the peers' slabs are mapped when world > 1: "ipc" (default; cudaMalloc + CUDA IPC
        handles over torch.distributed) or "torch" (the slab is allocated with torch.distributed._symmetric_memory --
        the allocator PyTorch backs with CUDA VMM / NVSHMEM -- and the peer pointers come from its rendezvous; this is
        the hook for externally managed symmetric heaps, reference bootstrap.cuh:359-360,442-443).  Environment
        override: FM_SYMM=ipc|torch."""
        self._L = _lib.load()
        self._ctx = ctypes.c_void_p()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device)
        self.rank, self.world = rank, world
        ccfg = _lib.FmConfig.from_config(cfg) if cfg is not None else None
        _lib.check(self._L.fm_create(ctypes.byref(ccfg) if ccfg is not None else None, rank, world, device,
                                     ctypes.byref(self._ctx)))
        d = _lib.FmDims()
        _lib.check(self._L.fm_get_dims(self._ctx, ctypes.byref(d)))
        self.dims: Dict[str, int] = {k: int(getattr(d, k)) for k, _ in d._fields_}
        self.cfg = cfg if cfg is not None else _lib.compiled_config()
        if timeout_ms is not None:
            _lib.check(self._L.fm_set_timeout_ms(self._ctx, int(timeout_ms)))
        self.symmetric = (symmetric or os.environ.get("FM_SYMM") or "ipc").lower()
        self._symm_tensor = None   # keeps an externally allocated slab alive
        if world > 1:
            if self.symmetric == "ipc":
                self._attach_peers(group)
            elif self.symmetric == "torch":
                self._attach_torch_symmetric_memory(group)
            else:
                raise ValueError(f"symmetric must be 'ipc' or 'torch', got {self.symmetric!r}")

    # ------------------------------------------------------------------ peers
    def _attach_peers(self, group) -> None:
        """Exchange CUDA IPC handles of the symmetric slab over torch.distributed and map every peer
        (replaces nvshmem_malloc + nvshmem_ptr, reference bootstrap.cuh:359-360,442-443)."""
        import torch.distributed as dist

        handle = (ctypes.c_ubyte * _lib.FM_IPC_HANDLE_BYTES)()
        _lib.check(self._L.fm_symm_export(self._ctx, handle))
        blob = exchange_blobs(bytes(handle), self.world, group)
        assert len(blob) == self.world * _lib.FM_IPC_HANDLE_BYTES
        _lib.check(self._L.fm_symm_attach_ipc(self._ctx, blob))
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)  # every rank's slab is mapped (and zeroed) before anyone dispatches into it

    def _attach_torch_symmetric_memory(self, group) -> None:
        """Externally allocated symmetric slab: torch.distributed._symmetric_memory.empty + rendezvous give every rank a
        same-sized allocation and the peers' mapped base pointers; the library adopts them through
        fm_symm_use_external + fm_symm_attach_ptrs (no cudaMalloc / CUDA IPC of its own)."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        nbytes = ctypes.c_size_t()
        _lib.check(self._L.fm_symm_size(self._ctx, ctypes.byref(nbytes)))
        n = (nbytes.value + 1023) // 1024 * 1024
        t = symm_mem.empty(n, dtype=torch.uint8, device=self.device)
        t.zero_()
        if t.data_ptr() % 1024:
            raise RuntimeError("symmetric memory allocation is not 1 KiB aligned")
        hdl = symm_mem.rendezvous(t, group if group is not None else dist.group.WORLD)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != t.data_ptr():
            raise RuntimeError("unexpected symmetric-memory rendezvous result")
        _lib.check(self._L.fm_symm_use_external(self._ctx, t.data_ptr(), n))
        arr = (ctypes.c_void_p * self.world)(*ptrs)
        _lib.check(self._L.fm_symm_attach_ptrs(self._ctx, arr))
        self._symm_tensor, self._symm_handle = t, hdl
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)

    # ------------------------------------------------------------------ hot path
    @property
    def num_local_experts(self) -> int:
        return self.dims["num_local_experts"]

    def _validate_weights(self, gate_weights: torch.Tensor, expert_weights: torch.Tensor) -> None:
        d = self.dims
        _check_tensor("Gate weights", gate_weights, self.device)
        _check_tensor("Expert weights", expert_weights, self.device)
        if gate_weights.dim() != 2 or gate_weights.size(0) != d["H"] or gate_weights.size(1) != d["E"]:
            raise RuntimeError(f"Gate weights must be [H={d['H']}, E={d['E']}]. Got {list(gate_weights.shape)}")
        nlx = d["num_local_experts"]
        if expert_weights.dim() != 4 or expert_weights.size(0) != nlx:
            raise RuntimeError(f"Expert count mismatch. Expected {nlx} local experts, got {expert_weights.size(0)}")
        if expert_weights.size(1) != 2:
            raise RuntimeError("Expert weights must have up and down projections [nLx, 2, P, H]")
        if expert_weights.size(2) != d["P"] or expert_weights.size(3) != d["H"]:
            raise RuntimeError(f"Expert weights must be [*, 2, P={d['P']}, H={d['H']}]. Got [*, 2, "
                               f"{expert_weights.size(2)}, {expert_weights.size(3)}]")

    def _validate(self, input: torch.Tensor, gate_weights: torch.Tensor, expert_weights: torch.Tensor) -> None:
        """Same conditions as the reference's TORCH_CHECKs (python_bindings.cu:22-65) plus dtype and alignment checks."""
        d = self.dims
        _check_tensor("Input", input, self.device)
        if input.dim() != 3:
            raise RuntimeError("Input must be 3D [batch, seq, H]")
        if input.size(0) * input.size(1) != d["S"]:
            raise RuntimeError(f"Input batch*seq must equal compiled S={d['S']}. Got batch={input.size(0)}, "
                               f"seq={input.size(1)} (product={input.size(0) * input.size(1)})")
        if input.size(2) != d["H"]:
            raise RuntimeError(f"Input hidden_size must equal compiled H={d['H']}. Got {input.size(2)}")
        self._validate_weights(gate_weights, expert_weights)

    def _bias_ptrs(self, bias_up, bias_down):
        d = self.dims
        nlx = d["num_local_experts"]
        for name, b, n in (("bias_up", bias_up, d["P"]), ("bias_down", bias_down, d["H"])):
            if b is not None:
                _check_tensor(name, b, self.device)
                if tuple(b.shape) != (nlx, n):
                    raise RuntimeError(f"{name} must be [{nlx}, {n}], got {list(b.shape)}")
        return (bias_up.data_ptr() if bias_up is not None else None,
                bias_down.data_ptr() if bias_down is not None else None)

    def forward(self, input: torch.Tensor, gate_weights: torch.Tensor, expert_weights: torch.Tensor,
                bias_up: Optional[torch.Tensor] = None, bias_down: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None, phase_mask: int = 7) -> torch.Tensor:
        """Enqueue one fused forward on torch's current stream; returns the [B,T,H] output tensor (async)."""
        self._validate(input, gate_weights, expert_weights)
        bu, bd = self._bias_ptrs(bias_up, bias_down)
        if out is None:
            out = torch.empty_like(input)
        else:
            _check_tensor("out", out, self.device)
            if out.shape != input.shape:
                raise RuntimeError("out must have the input's shape")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if phase_mask == 7:
            rc = self._L.fm_moe_forward(self._ctx, input.data_ptr(), gate_weights.data_ptr(),
                                        expert_weights.data_ptr(), bu, bd, out.data_ptr(), stream)
        else:
            rc = self._L.fm_debug_forward(self._ctx, input.data_ptr(), gate_weights.data_ptr(),
                                          expert_weights.data_ptr(), bu, bd, out.data_ptr(), stream, phase_mask)
        _lib.check(rc)
        return out

    def _host_args(self, input_host, gate_weights, expert_weights, out_host, bias_up, bias_down):
        d = self.dims
        for name, t in (("input_host", input_host), ("out_host", out_host)):
            if t is None:
                continue
            if t.is_cuda or t.dtype != torch.bfloat16 or not t.is_contiguous():
                raise RuntimeError(f"{name} must be a contiguous CPU torch.bfloat16 tensor")
            if t.numel() != d["S"] * d["H"]:
                raise RuntimeError(f"{name} must hold S*H = {d['S'] * d['H']} elements")
        self._validate_weights(gate_weights, expert_weights)
        bu, bd = self._bias_ptrs(bias_up, bias_down)
        if out_host is None:
            out_host = torch.empty_like(input_host, pin_memory=True)
        return out_host, bu, bd

    def forward_host(self, input_host: torch.Tensor, gate_weights: torch.Tensor, expert_weights: torch.Tensor,
                     out_host: Optional[torch.Tensor] = None, bias_up: Optional[torch.Tensor] = None,
                     bias_down: Optional[torch.Tensor] = None) -> torch.Tensor:
        """End-to-end call with HOST activations (pinned for full PCIe rate): H2D copy, fused forward, D2H copy, wait.
        Weights stay device-resident (they are parameters, not per-step inputs).  Blocking; for throughput use
        submit_host / wait_host, which keep up to three steps in flight so the copies overlap the kernels."""
        out_host, bu, bd = self._host_args(input_host, gate_weights, expert_weights, out_host, bias_up, bias_down)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._L.fm_moe_forward_host(self._ctx, input_host.data_ptr(), gate_weights.data_ptr(),
                                               expert_weights.data_ptr(), bu, bd, out_host.data_ptr(), stream))
        return out_host

    def submit_host(self, input_host: torch.Tensor, gate_weights: torch.Tensor, expert_weights: torch.Tensor,
                    out_host: torch.Tensor, bias_up: Optional[torch.Tensor] = None,
                    bias_down: Optional[torch.Tensor] = None) -> int:
        """Asynchronous host-buffer step: enqueue H2D copy -> layer -> D2H copy and return a ticket at once
        (at most `_lib.FM_HOST_SLOTS` tickets outstanding).  `out_host` is complete after `wait_host(ticket)`."""
        out_host, bu, bd = self._host_args(input_host, gate_weights, expert_weights, out_host, bias_up, bias_down)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ticket = ctypes.c_uint64()
        _lib.check(self._L.fm_host_submit(self._ctx, input_host.data_ptr(), gate_weights.data_ptr(),
                                          expert_weights.data_ptr(), bu, bd, out_host.data_ptr(), stream,
                                          ctypes.byref(ticket)))
        return int(ticket.value)

    def wait_host(self, ticket: int) -> None:
        _lib.check(self._L.fm_host_wait(self._ctx, int(ticket)))

    def output_buffer(self) -> torch.Tensor:
        """A [mini_batch, seq, H] bf16 tensor to pass as `out=`.  With world > 1 it aliases the region of this rank's
        symmetric slab that the experts accumulate into, so `forward(..., out=ctx.output_buffer())` leaves the result in
        place instead of copying it to a caller tensor at the end of the kernel (it is overwritten by the next forward
        and dies with the context).  With world == 1 the kernel accumulates into any `out`; a fresh tensor is returned."""
        shape = (self.cfg.mini_batch, self.cfg.sequence_len, self.dims["H"])
        ptr, nbytes = ctypes.c_void_p(), ctypes.c_size_t()
        _lib.check(self._L.fm_output_buffer(self._ctx, ctypes.byref(ptr), ctypes.byref(nbytes)))
        if not ptr.value:
            return torch.empty(shape, dtype=torch.bfloat16, device=self.device)

        class _Raw:  # CUDA array interface view of the slab region (int16 elements, reinterpreted below)
            __cuda_array_interface__ = {"shape": (nbytes.value // 2,), "typestr": "<i2", "data": (ptr.value, False),
                                        "version": 3}

        return torch.as_tensor(_Raw(), device=self.device).view(torch.bfloat16).view(shape)

    def sync_ranks(self, group=None) -> None:
        """Optional host-side rendezvous before a forward (world > 1).  The kernel's waits are bounded (default 120 s,
        `timeout_ms`) and end in a trap when a peer never shows up; ranks whose skew may exceed that -- first-iteration
        compilation, checkpoint I/O, a debugger -- can line up here first."""
        if self.world > 1:
            import torch.distributed as dist

            torch.cuda.synchronize(self.device)
            dist.barrier(group=group)

    def set_trace(self, enable: bool) -> None:
        _lib.check(self._L.fm_set_trace(self._ctx, 1 if enable else 0))

    def check(self) -> None:
        """Raise if the last kernel reported a protocol timeout (call after a synchronise)."""
        _lib.check(self._L.fm_check(self._ctx))

    def synchronize(self) -> None:
        try:
            torch.cuda.synchronize(self.device)
        finally:
            self.check()

    def read(self, name: str) -> np.ndarray:
        """Copy an internal buffer (routing tables, staging areas) to host -- tests and debugging only."""
        which, dt = _BUFFER_SPECS[name]
        nbytes = ctypes.c_size_t()
        _lib.check(self._L.fm_buffer_bytes(self._ctx, which, ctypes.byref(nbytes)))
        arr = np.empty(nbytes.value // np.dtype(dt).itemsize, dtype=dt)
        _lib.check(self._L.fm_read_buffer(self._ctx, which, arr.ctypes.data_as(ctypes.c_void_p), nbytes.value))
        d = self.dims
        k, E, S, H, P = d["k"], d["E"], d["S"], d["H"], d["P"]
        npk = d["world"] * d["num_local_experts"]
        shapes = {"topk_idx": (S, k), "topk_w": (S, k), "mcw": (S,), "slot": (S, k), "counts": (E,),
                  "recv_x": (npk, d["pEC"], H), "hidden": (npk, d["pEC"], P), "ret_y": (E, d["pEC"], H),
                  "gate_out": (S, E), "recv_cnt": (npk,), "trace": (d["grid"], 128), "aux_loss": (2 * E + 1,)}
        return arr.reshape(shapes[name])

    @property
    def launch_count(self) -> int:
        return int(self._L.fm_launch_count(self._ctx))

    def close(self) -> None:
        if self._ctx:
            self._L.fm_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
