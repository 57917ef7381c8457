"""Generates the committed golden fixtures tests/golden/*.npz.

PARITY UNPINNED: the reference ships no golden vectors for this path and cannot run in this environment, so these
fixtures are produced by the oracle (oracle/moe_oracle.c) on seeded inputs.  They pin the ORACLE against
regressions and let the GPU box (where /root/reference does not exist) check the CUDA path against fixed files; they
do not pin the oracle to the reference.  Inputs are regenerated from the seed inside the tests; the fixtures store the
seed, the configuration and the oracle's outputs.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from flashmoe_b200.config import MoEConfig  # noqa: E402
from tests.util import make_inputs, run_oracle  # noqa: E402

CASES = {
    # BASELINE.json configs[0]: "2 experts, top-1, seq=128, d_model=512" (ffn unspecified -> 4*H)
    "configA_k1": dict(cfg=MoEConfig(num_experts=2, expert_top_k=1, sequence_len=128, hidden_size=512,
                                     intermediate_size=2048), seed=11, scaled=True, bias=False),
    "tiny_k2_drop": dict(cfg=MoEConfig(num_experts=4, expert_top_k=2, sequence_len=128, hidden_size=64,
                                       intermediate_size=256), seed=12, scaled=True, bias=False),
    "small_k2_nodrop_gelu_bias": dict(cfg=MoEConfig(num_experts=8, expert_top_k=2, sequence_len=256, hidden_size=128,
                                                    intermediate_size=320, drop_tokens=0, hidden_act=1), seed=13,
                                      scaled=True, bias=True),
    "unscaled_underflow_k3": dict(cfg=MoEConfig(num_experts=8, expert_top_k=3, sequence_len=128, hidden_size=256,
                                                intermediate_size=128), seed=14, scaled=False, bias=False),
}


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, c in CASES.items():
        cfg = c["cfg"]
        x, wg, we, bu, bd = make_inputs(cfg, seed=c["seed"], scaled=c["scaled"], bias=c["bias"])
        r = run_oracle(cfg, x, wg, we, bu, bd)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), config=np.array([cfg.raw()[k] for k in sorted(cfg.raw())]),
                            config_keys=np.array(sorted(cfg.raw())), seed=c["seed"], scaled=c["scaled"], bias=c["bias"],
                            out=r.out, topk_idx=r.topk_idx, slot=r.slot, kept=r.kept, counts=r.counts, mcw=r.mcw,
                            gate_out=r.gate_out, ambiguous=r.ambiguous)
        print(name, "S", cfg.S, "dropped", int((r.kept == 0).sum()), "ambiguous", int(r.ambiguous.sum()))


if __name__ == "__main__":
    main()
