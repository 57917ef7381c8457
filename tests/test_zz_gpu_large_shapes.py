"""BASELINE.json shapes beyond config B (d_model 2048 / 4096): the router's multi-piece GEMV path (H > 1024 columns per
staged chunk) and the dispatch's multi-group row staging (chunk rows > 128 KiB of shared memory) only occur at these sizes.
Runs last (file name) because each case needs a few seconds of CPU oracle time."""
import pytest

from flashmoe_b200.config import BASELINE_CONFIGS, MoEConfig
from tests.test_gpu_parity import _compare, _run
from tests.util import make_inputs, run_oracle

pytestmark = pytest.mark.gpu


def test_config_E8_full_size_parity():
    """expert-sweep point E=8: 8 experts, top-2, seq 8192, d_model 2048, ffn 2048."""
    cfg = BASELINE_CONFIGS["E8"]
    x, wg, we, _, _ = make_inputs(cfg, seed=41)
    # fp32 summation order over 2048-4096 terms moves the logits by ~1e-6: looser bound on the probability sum
    _compare(cfg, _run(cfg, x, wg, we), run_oracle(cfg, x, wg, we), mcw_rtol=2e-5)


def test_config_C_hidden_4096_reduced_ffn_parity():
    """config C's d_model (4096) and token count with a reduced ffn (1024) so the oracle stays at seconds."""
    cfg = MoEConfig(num_experts=8, expert_top_k=2, sequence_len=4096, hidden_size=4096, intermediate_size=1024)
    x, wg, we, _, _ = make_inputs(cfg, seed=42)
    _compare(cfg, _run(cfg, x, wg, we), run_oracle(cfg, x, wg, we), mcw_rtol=2e-5)


def test_token_sweep_point_32_experts():
    """config D shape: 32 experts, top-2, d_model 2048, ffn 2048 at 4096 tokens (capacity 256 rows per expert)."""
    cfg = BASELINE_CONFIGS["D4k"]
    x, wg, we, _, _ = make_inputs(cfg, seed=43)
    _compare(cfg, _run(cfg, x, wg, we), run_oracle(cfg, x, wg, we), mcw_rtol=2e-5)


def test_module_wrapper_with_bias_and_routing_outputs():
    """flashmoe_b200.layer.FlashMoELayer: caller-owned parameters, bias + GELU, routing tables returned."""
    import torch

    from flashmoe_b200.layer import FlashMoELayer
    from oracle import moe_oracle as mo
    from tests.util import check_output, check_topk

    cfg = MoEConfig(num_experts=8, expert_top_k=2, sequence_len=256, hidden_size=256, intermediate_size=512, hidden_act=1)
    torch.manual_seed(5)
    layer = FlashMoELayer(cfg, bias=True)
    layer.bias_up.data.normal_(0, 0.1)
    layer.bias_down.data.normal_(0, 0.1)
    x = torch.randn(1, cfg.S, cfg.H, device=layer.ctx.device).bfloat16()
    res = layer(x, return_routing=True)
    assert res.out.shape == x.shape and res.topk_idx.shape == (cfg.S, 2) and res.topk_weight.dtype == torch.bfloat16
    ref = run_oracle(cfg, x.cpu(), layer.gate_weight.data.cpu(), layer.expert_weight.data.cpu(), layer.bias_up.data.cpu(),
                     layer.bias_down.data.cpu())
    mism = check_topk(res.topk_idx.numpy(), ref)
    check_output(mo.to_bits(res.out.cpu().reshape(cfg.S, cfg.H)), ref.out, rows_ok=~mism)
    assert (res.slot.numpy()[~mism] == ref.slot[~mism]).all()
    layer.ctx.close()


def test_run_moe_entry_point_single_gpu():
    """flashmoe.run_moe(): launcher -> worker -> _C on the compiled configuration, like the reference's quick start."""
    import flashmoe

    res = flashmoe.run_moe(n_processes=1)
    assert "FlashMoE forward pass took" in res.stdout and "Completed! Output: torch.Size([1, 4096, 1024])" in res.stdout
