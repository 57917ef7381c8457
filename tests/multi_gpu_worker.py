"""Launched by torchrun (one rank per GPU): expert-parallel forward over NVLink peer memory, each rank checks its own
[S,H] output against the oracle's world composition.  Prints 'RANK r PARITY OK'."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from flashmoe_b200.config import MoEConfig  # noqa: E402
from flashmoe_b200.runtime import MoEContext, env_rank_world  # noqa: E402
from oracle import moe_oracle as mo  # noqa: E402
from tests.util import check_output, check_topk  # noqa: E402


def main():
    rank, world, local = env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world, device_id=dev)
    cases = json.loads(os.environ.get("FM_MULTI_CASES", "null")) or [
        dict(num_experts=8, expert_top_k=2, sequence_len=512, hidden_size=256, intermediate_size=512),
        dict(num_experts=8, expert_top_k=2, sequence_len=1024, hidden_size=512, intermediate_size=1024, drop_tokens=0),
        dict(num_experts=16, expert_top_k=4, sequence_len=256, hidden_size=128, intermediate_size=256, hidden_act=1),
    ]
    for ci, kw in enumerate(cases):
        cfg = MoEConfig(**kw)
        if cfg.E % world:
            continue
        nlx = cfg.num_local_experts(world)
        g = torch.Generator().manual_seed(1000 + 17 * ci + rank)
        gw = torch.Generator().manual_seed(777 + ci)
        x = torch.randn(1, cfg.S, cfg.H, generator=g).bfloat16()
        wg = (torch.randn(cfg.H, cfg.E, generator=gw) * cfg.H ** -0.5).bfloat16()
        we = (torch.randn(nlx, 2, cfg.P, cfg.H, generator=g) * cfg.H ** -0.5).bfloat16()
        ctx = MoEContext(cfg, rank=rank, world=world, device=local, timeout_ms=20000)
        xd, wgd, wed = x.to(dev), wg.to(dev), we.to(dev)
        out = None
        for _ in range(int(os.environ.get("FM_MULTI_LAUNCHES", "3"))):  # several launches: epoch-tagged flags, buffer reuse
            out = ctx.forward(xd, wgd, wed)
        ctx.synchronize()
        dist.barrier()
        # gather everything the oracle needs (all ranks' tokens are irrelevant to mine; all experts are relevant)
        wes = [torch.empty_like(wed) for _ in range(world)]
        dist.all_gather(wes, wed)
        full = torch.cat([w.cpu() for w in wes], dim=0)
        up, down = mo.split_expert_weights(mo.to_bits(full))
        ref = mo.forward(mo.to_bits(x.view(cfg.S, cfg.H)), mo.gate_weights_effective(mo.to_bits(wg), cfg.E, cfg.H), up,
                         down, k=cfg.k, EC=cfg.EC, act=cfg.hidden_act)
        mism = check_topk(ctx.read("topk_idx"), ref)
        relf = check_output(mo.to_bits(out.cpu().view(cfg.S, cfg.H)), ref.out, rows_ok=~mism)
        recv = ctx.read("recv_cnt")
        print(f"RANK {rank} case {ci} PARITY OK relF={relf:.2e} recv_cnt={recv.tolist()}", flush=True)
        dist.barrier()
        ctx.close()
        dist.barrier()
    dist.destroy_process_group()
    print(f"RANK {rank} ALL OK", flush=True)


if __name__ == "__main__":
    main()
