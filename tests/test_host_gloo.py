"""Host-side multi-rank logic on CPU with the gloo backend (world_size 2): rank/world discovery from the launcher
environment, the out-of-band exchange that carries the CUDA IPC handles, per-rank synthetic inputs (gate weights
identical on all ranks, experts sharded) and the oracle's world composition computed rank-locally."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flashmoe_b200.config import MoEConfig


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from flashmoe_b200 import worker as W
    from flashmoe_b200.runtime import env_rank_world, exchange_blobs
    from oracle import moe_oracle as mo

    assert env_rank_world() == (rank, world, rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # 1. the handle exchange: every rank ends with all handles in rank order
    mine = bytes([rank]) * 64
    blob = exchange_blobs(mine, world, None)
    assert blob == b"".join(bytes([r]) * 64 for r in range(world))
    # 2. sharded synthetic inputs + rank-local oracle over all-gathered expert weights
    cfg = MoEConfig(num_experts=4, expert_top_k=2, sequence_len=128, hidden_size=64, intermediate_size=128)
    nlx = cfg.num_local_experts(world)
    x, wg, we = W.make_inputs(cfg, nlx, rank, "cpu", scaled=True)
    gathered = [torch.empty_like(we.view(torch.uint8)) for _ in range(world)]
    dist.all_gather(gathered, we.view(torch.uint8).contiguous())
    wgs = [torch.empty_like(wg.view(torch.uint8)) for _ in range(world)]
    dist.all_gather(wgs, wg.view(torch.uint8).contiguous())
    assert all((w == wgs[0]).all() for w in wgs), "gate weights must be identical on every rank"
    full = torch.cat([g.view(torch.bfloat16) for g in gathered], dim=0)
    up, down = mo.split_expert_weights(mo.to_bits(full))
    r = mo.forward(mo.to_bits(x.view(cfg.S, cfg.H)), mo.gate_weights_effective(mo.to_bits(wg), cfg.E, cfg.H), up, down,
                   k=cfg.k, EC=cfg.EC)
    q.put((rank, r.out.copy(), r.topk_idx.copy(), mo.to_bits(x.view(cfg.S, cfg.H)), mo.to_bits(wg), mo.to_bits(we)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_host_logic_with_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import moe_oracle as mo

    res = mo.forward_world([g[3] for g in got], [g[4] for g in got], [g[5] for g in got], k=2, EC=64)
    for r in range(world):
        assert (got[r][1] == res[r].out).all() and (got[r][2] == res[r].topk_idx).all()
    assert not (got[0][3] == got[1][3]).all(), "ranks must own different tokens"


def test_env_rank_world_fallbacks(monkeypatch):
    from flashmoe_b200.runtime import env_rank_world

    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE", "PMI_RANK", "PMI_SIZE",
              "SLURM_PROCID", "SLURM_NTASKS", "OMPI_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID"):
        monkeypatch.delenv(k, raising=False)
    assert env_rank_world() == (0, 1, 0)
    monkeypatch.setenv("OMPI_COMM_WORLD_RANK", "3")
    monkeypatch.setenv("OMPI_COMM_WORLD_SIZE", "8")
    assert env_rank_world() == (3, 8, 3)
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("LOCAL_RANK", "1")
    assert env_rank_world() == (1, 2, 1)
