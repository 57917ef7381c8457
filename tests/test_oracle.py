"""The CPU oracle against the committed golden fixtures, hand-made corner cases of SURVEY.md Appendix A and an
independent plain-torch implementation (oracle/torch_moe.py).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from flashmoe_b200.config import MoEConfig
from oracle import moe_oracle as mo
from oracle import torch_moe as tm
from tests.util import check_output, make_inputs, run_oracle

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def load_golden(path):
    z = np.load(path, allow_pickle=False)
    cfg = MoEConfig(**{str(k): int(v) for k, v in zip(z["config_keys"], z["config"])})
    return cfg, z


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    cfg, z = load_golden(path)
    x, wg, we, bu, bd = make_inputs(cfg, seed=int(z["seed"]), scaled=bool(z["scaled"]), bias=bool(z["bias"]))
    r = run_oracle(cfg, x, wg, we, bu, bd)
    assert (r.topk_idx == z["topk_idx"]).all() and (r.slot == z["slot"]).all() and (r.counts == z["counts"]).all()
    assert (r.gate_out == z["gate_out"]).all() and (r.mcw == z["mcw"]).all()
    assert (r.out == z["out"]).all()  # bit-exact: same code, same seeded inputs


def test_golden_set_is_present():
    assert len(GOLDEN) >= 4


def _bits(t):
    return mo.to_bits(t.bfloat16())


def test_ties_resolve_to_lowest_expert_index():
    # experts 1 and 3 share a gate row -> identical logits and probabilities -> strict '>' keeps the lower index
    H, E = 64, 4
    g = torch.Generator().manual_seed(0)
    x = torch.randn(32, H, generator=g).bfloat16()
    w = torch.randn(E, H, generator=g) * 0.1
    w[3] = w[1]
    w[0] = -w[1]
    w[2] = w[0]
    up = _bits(torch.randn(E, 128, H, generator=g) * 0.1)
    down = _bits(torch.randn(E, H, 128, generator=g) * 0.1)
    r = mo.forward(mo.to_bits(x), _bits(w), up, down, k=2, EC=64)
    for t in range(32):
        a, b = r.topk_idx[t]
        assert (a, b) in ((1, 3), (0, 2)), (a, b)


def test_exp_underflow_picks_lowest_unselected_index_not_largest_logit():
    # Appendix A.4: logits [100, -200, -150, -120]: p = [1, 0, 0, 0] after the ftz flush, so picks are 0, 1, 2 in
    # index order although expert 3 has the largest remaining logit; mCw = 1
    H, E = 64, 4
    x = torch.zeros(8, H)
    x[:, 0] = 1.0
    w = torch.zeros(E, H)
    w[:, 0] = torch.tensor([100.0, -200.0, -150.0, -120.0])
    g = torch.Generator().manual_seed(1)
    up = _bits(torch.randn(E, 64, H, generator=g))
    down = _bits(torch.randn(E, H, 64, generator=g))
    r = mo.forward(_bits(x), _bits(w), up, down, k=3, EC=8)
    assert (r.topk_idx == np.array([0, 1, 2])).all()
    assert np.allclose(r.mcw, 1.0)
    # the zero-probability picks still consume capacity slots (counts include them) ...
    assert r.counts.tolist() == [8, 8, 8, 0]
    # ... and contribute 0 * y: the output equals expert 0's y scaled by p~=1, mCw=1
    h, y = mo.expert_ffn(_bits(x), up[0], down[0])
    assert (r.out == y).all()


def test_capacity_drops_in_ascending_token_order_and_dropped_pairs_contribute_nothing():
    H, E, S = 64, 4, 128
    cfg = MoEConfig(num_experts=E, expert_top_k=1, sequence_len=S, hidden_size=H, intermediate_size=64)
    x = torch.ones(S, H) * 0.5
    w = torch.zeros(E, H)
    w[2] = 1.0  # every token picks expert 2
    g = torch.Generator().manual_seed(2)
    up = _bits(torch.randn(E, 64, H, generator=g) * 0.2)
    down = _bits(torch.randn(E, H, 64, generator=g) * 0.2)
    r = mo.forward(_bits(x), _bits(w), up, down, k=1, EC=cfg.EC)
    assert cfg.EC == 32
    assert (r.topk_idx[:, 0] == 2).all() and r.counts.tolist() == [0, 0, S, 0]
    assert (r.slot[:, 0] == np.arange(S)).all()
    assert r.kept[:32].all() and not r.kept[32:].any()
    assert (r.out[32:] == 0).all() and (r.out[:32] != 0).any()


def test_top1_combine_is_unscaled_copy():
    cfg = MoEConfig(num_experts=2, expert_top_k=1, sequence_len=128, hidden_size=64, intermediate_size=128, drop_tokens=0)
    x, wg, we, _, _ = make_inputs(cfg, seed=3)
    r = run_oracle(cfg, x, wg, we)
    up, down = mo.split_expert_weights(mo.to_bits(we))
    xb = mo.to_bits(x.reshape(cfg.S, cfg.H))
    for t in (0, 17, 127):
        e = int(r.topk_idx[t, 0])
        _, y = mo.expert_ffn(xb[t:t + 1], up[e], down[e])
        assert (r.out[t] == y[0]).all()  # no gate scaling for k == 1 (processor.cuh:170-203)


def test_gate_and_down_weights_are_reinterpreted_not_transposed():
    H, E, P = 64, 4, 128
    g = torch.Generator().manual_seed(4)
    wg = torch.randn(H, E, generator=g).bfloat16()
    eff = mo.gate_weights_effective(mo.to_bits(wg), E, H)
    assert (eff.reshape(-1) == mo.to_bits(wg).reshape(-1)).all() and eff.shape == (E, H)
    assert not (eff == mo.to_bits(wg.t().contiguous())).all()
    we = torch.randn(2, 2, P, H, generator=g).bfloat16()
    up, down = mo.split_expert_weights(mo.to_bits(we))
    assert up.shape == (2, P, H) and down.shape == (2, H, P)
    assert (down[1].reshape(-1) == mo.to_bits(we)[1, 1].reshape(-1)).all()


@pytest.mark.parametrize("E,k,drop,act", [(8, 2, 1, 0), (8, 2, 0, 1), (16, 4, 1, 0), (2, 1, 1, 0)])
def test_c_oracle_agrees_with_plain_torch_forward(E, k, drop, act):
    cfg = MoEConfig(num_experts=E, expert_top_k=k, sequence_len=256, hidden_size=128, intermediate_size=256,
                    drop_tokens=drop, hidden_act=act)
    x, wg, we, _, _ = make_inputs(cfg, seed=5)
    r = run_oracle(cfg, x, wg, we)
    t = tm.moe_forward_cpu(x.reshape(cfg.S, cfg.H), wg, we, k=k, EC=cfg.EC, act=act)
    mism = (r.topk_idx != t["topk_idx"].numpy()).any(axis=1)
    assert not (mism & ~r.ambiguous).any()
    ok_rows = ~mism
    assert (r.slot[ok_rows] == t["slot"].numpy()[ok_rows]).all() or mism.any()
    check_output(mo.to_bits(t["out"]), r.out, rows_ok=ok_rows, what="torch vs C oracle")


def test_world_composition_matches_single_rank_view():
    # Appendix A.8: each rank's tokens see an independent problem over the concatenation of all ranks' experts
    W, nlx = 2, 2
    cfg = MoEConfig(num_experts=W * nlx, expert_top_k=2, sequence_len=128, hidden_size=64, intermediate_size=128)
    xs, wes = [], []
    g = torch.Generator().manual_seed(6)
    wg = (torch.randn(cfg.H, cfg.E, generator=g) * cfg.H ** -0.5).bfloat16()
    for r in range(W):
        xs.append(torch.randn(cfg.S, cfg.H, generator=g).bfloat16())
        wes.append((torch.randn(nlx, 2, cfg.P, cfg.H, generator=g) * cfg.H ** -0.5).bfloat16())
    res = mo.forward_world([mo.to_bits(x) for x in xs], [mo.to_bits(wg)] * W, [mo.to_bits(w) for w in wes], k=cfg.k,
                           EC=cfg.EC)
    full = torch.cat(wes, dim=0)
    for r in range(W):
        single = run_oracle(cfg, xs[r], wg, full)
        assert (single.out == res[r].out).all() and (single.topk_idx == res[r].topk_idx).all()


def test_ambiguity_flags_fire_on_near_ties_only():
    logits = np.array([[1.0, 0.5, 0.49999, -3.0], [1.0, 0.5, 0.0, -3.0]], dtype=np.float32)
    flags = mo.ambiguity_flags(logits, np.array([100.0, 100.0], dtype=np.float32), k=2)
    assert flags.tolist() == [True, False]


def test_aux_loss_restatement_against_a_numpy_computation():
    """is_training = 1: gML = mean_t p[t,e], gMeC = counts[e] / S, loss = sum gML*gMeC / E (gate.cuh:608-635,698-706,763-773)."""
    cfg = MoEConfig(num_experts=8, expert_top_k=2, sequence_len=256, hidden_size=128, intermediate_size=128)
    x, wg, we, _, _ = make_inputs(cfg, seed=11)
    ref = run_oracle(cfg, x, wg, we)
    gml, gmec, loss = mo.aux_loss(mo.to_bits(x.reshape(cfg.S, cfg.H)), mo.gate_weights_effective(mo.to_bits(wg), cfg.E, cfg.H),
                                  k=cfg.k, EC=cfg.EC)
    # independent: library softmax of the oracle's logits in float64
    l = ref.logits.astype(np.float64)
    p = np.exp(l - l.max(axis=1, keepdims=True))
    p /= p.sum(axis=1, keepdims=True)
    np.testing.assert_allclose(gml, p.mean(axis=0), rtol=1e-5)
    np.testing.assert_allclose(gmec, ref.counts / cfg.S, rtol=1e-7)
    np.testing.assert_allclose(loss, float((p.mean(axis=0) * ref.counts / cfg.S).sum() / cfg.E), rtol=1e-5)
    assert abs(gml.sum() - 1.0) < 1e-5 and abs(gmec.sum() - cfg.k) < 1e-6


def test_sampled_forward_equals_full_forward_on_the_sample():
    """fmo_forward_sample (bench.py's post-run parity check, full-size multi-GPU parity): routing / slots / drops over all
    tokens, FFN + combine only for the sample -- bit-identical to the full forward on those tokens, duplicates allowed."""
    cfg = MoEConfig(num_experts=8, expert_top_k=2, sequence_len=512, hidden_size=256, intermediate_size=512)
    x, wg, we, _, _ = make_inputs(cfg, seed=3)
    ref = run_oracle(cfg, x, wg, we)
    up, down = mo.split_expert_weights(mo.to_bits(we))
    sample = np.array([5, 0, 511, 77, 300, 300], dtype=np.int32)
    r = mo.forward_sample(mo.to_bits(x.reshape(cfg.S, cfg.H)), mo.gate_weights_effective(mo.to_bits(wg), cfg.E, cfg.H), up, down,
                          sample, k=cfg.k, EC=cfg.EC)
    assert (r.out == ref.out[sample]).all()
    assert (r.topk_idx == ref.topk_idx).all() and (r.slot == ref.slot).all() and (r.counts == ref.counts).all()
    assert (r.kept == ref.kept).all() and (r.ambiguous == ref.ambiguous).all()
