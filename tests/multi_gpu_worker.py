"""Launched by torchrun (one rank per GPU): expert-parallel forward over NVLink peer memory, each rank checks its own
[S,H] output against the oracle's world composition.  Prints 'RANK r ALL OK'.

Every case runs several back-to-back launches on ONE context with DIFFERENT activations per launch (routing, per-packet
row counts and drops change from launch to launch, so a GEMM that read a stale row of the previous epoch -- recv_x,
hidden, recv_meta, the accumulator -- would produce a wrong result), and EVERY launch's output is compared with the
oracle.  Full-size cases (config B sharded, the 32-expert token-sweep shape) are checked on a seeded sample of tokens per
rank (the oracle's FFN cost is per token; routing and capacity drops are still computed over all tokens)."""
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

DEFAULT_CASES = [
    dict(num_experts=8, expert_top_k=2, sequence_len=512, hidden_size=256, intermediate_size=512),
    dict(num_experts=8, expert_top_k=2, sequence_len=1024, hidden_size=512, intermediate_size=1024, drop_tokens=0),
    dict(num_experts=16, expert_top_k=4, sequence_len=256, hidden_size=128, intermediate_size=256, hidden_act=1),
    # BASELINE config B sharded E/W -- the shape bench.py times at N = 2, 4, 8 (sampled oracle)
    dict(num_experts=8, expert_top_k=2, sequence_len=4096, hidden_size=1024, intermediate_size=4096, _sample=192,
         _launches=20),
    # token-sweep shape D4k: 32 experts, d_model 2048, ffn 2048 (sampled oracle)
    dict(num_experts=32, expert_top_k=2, sequence_len=4096, hidden_size=2048, intermediate_size=2048, _sample=128),
]


def main():
    rank, world, local = env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world, device_id=dev)
    cases = json.loads(os.environ.get("FM_MULTI_CASES", "null")) or DEFAULT_CASES
    n_inputs = 3
    for ci, kw in enumerate(cases):
        kw = dict(kw)
        n_sample = int(kw.pop("_sample", 0))
        launches = int(kw.pop("_launches", os.environ.get("FM_MULTI_LAUNCHES", "6")))
        cfg = MoEConfig(**kw)
        if cfg.E % world:
            continue
        nlx = cfg.num_local_experts(world)
        g = torch.Generator().manual_seed(1000 + 17 * ci + rank)
        gw = torch.Generator().manual_seed(777 + ci)
        xs = [torch.randn(1, cfg.S, cfg.H, generator=g).bfloat16() for _ in range(n_inputs)]
        # the third input routes most tokens to few experts: counts shrink / grow strongly between launches.  (Token 0 plus
        # small noise, not identical copies: with identical rows one borderline element would be replicated S/2 times and
        # decide the ">= 99.9 % of the elements within tolerance" criterion on its own.)
        noise = torch.randn(1, cfg.S // 2, cfg.H, generator=g) * 0.02
        xs[2][:, : cfg.S // 2] = (xs[2][:, :1].float() + noise).bfloat16()
        wg = (torch.randn(cfg.H, cfg.E, generator=gw) * cfg.H ** -0.5).bfloat16()
        we = (torch.randn(nlx, 2, cfg.P, cfg.H, generator=g) * cfg.H ** -0.5).bfloat16()
        ctx = MoEContext(cfg, rank=rank, world=world, device=local, timeout_ms=20000)
        xds, wgd, wed = [x.to(dev) for x in xs], wg.to(dev), we.to(dev)
        # everything the oracle needs: all experts' weights (other ranks' tokens are irrelevant to mine)
        wes = [torch.empty_like(wed) for _ in range(world)]
        dist.all_gather(wes, wed)
        up, down = mo.split_expert_weights(mo.to_bits(torch.cat([w.cpu() for w in wes], dim=0)))
        del wes
        wg_eff = mo.gate_weights_effective(mo.to_bits(wg), cfg.E, cfg.H)
        refs, samples = [], []
        for i, x in enumerate(xs):
            xb = mo.to_bits(x.view(cfg.S, cfg.H))
            if n_sample:
                smp = np.sort(np.random.default_rng(5 + i + rank).choice(cfg.S, size=n_sample, replace=False)).astype(np.int32)
                refs.append(mo.forward_sample(xb, wg_eff, up, down, smp, k=cfg.k, EC=cfg.EC, act=cfg.hidden_act))
            else:
                smp = None
                refs.append(mo.forward(xb, wg_eff, up, down, k=cfg.k, EC=cfg.EC, act=cfg.hidden_act))
            samples.append(smp)
        out_slab = ctx.output_buffer()
        worst = 0.0
        for it in range(launches):
            i = it % n_inputs
            # alternate between a caller tensor and the zero-copy slab output
            out = ctx.forward(xds[i], wgd, wed, out=out_slab if it % 2 else None)
            ctx.synchronize()
            ref, smp = refs[i], samples[i]
            mism = check_topk(ctx.read("topk_idx"), ref)
            got = mo.to_bits(out.cpu().view(cfg.S, cfg.H))
            if smp is not None:
                relf = check_output(got[smp], ref.out, rows_ok=~mism[smp], what=f"rank {rank} case {ci} launch {it}")
            else:
                relf = check_output(got, ref.out, rows_ok=~mism, what=f"rank {rank} case {ci} launch {it}")
                if not mism.any():
                    assert (ctx.read("slot") == ref.slot).all() and (ctx.read("counts") == ref.counts).all()
            worst = max(worst, relf)
        recv = ctx.read("recv_cnt")
        print(f"RANK {rank} case {ci} PARITY OK launches={launches} relF_max={worst:.2e} recv_cnt={recv.tolist()[:16]}", flush=True)
        dist.barrier()
        ctx.close()
        dist.barrier()
    dist.destroy_process_group()
    print(f"RANK {rank} ALL OK", flush=True)


if __name__ == "__main__":
    main()
