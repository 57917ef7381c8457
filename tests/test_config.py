"""Compile-time configuration surface (reference csrc/flashmoe_config.json + schema, types.cuh:441-512)."""
import json

import pytest

from flashmoe_b200 import config as C


def test_default_json_is_valid_and_is_config_B():
    cfg = C.load_config()
    cfg.check_hot_path()
    assert (cfg.S, cfg.H, cfg.P, cfg.E, cfg.k) == (4096, 1024, 4096, 8, 2)
    assert cfg.torch_dtype == C.DTYPE_BF16
    assert set(cfg.raw()) == set(C.ALL_KEYS) and len(C.ALL_KEYS) == 15


@pytest.mark.parametrize("name,S,PX,EC,pEC,TCM", [
    # SURVEY.md Appendix B table (drop_tokens=1, capacity_factor=1)
    ("A", 128, 64, 64, 128, 1),
    ("B", 4096, 64, 1024, 1024, 8),
    ("C", 4096, 64, 1024, 1024, 8),
    ("D4k", 4096, 64, 256, 256, 2),
    ("E8", 8192, 64, 2048, 2048, 16),
])
def test_derived_constants_match_reference_acc(name, S, PX, EC, pEC, TCM):
    cfg = C.BASELINE_CONFIGS[name]
    assert (cfg.S, cfg.PX, cfg.EC, cfg.pEC, cfg.TCM) == (S, PX, EC, pEC, TCM)


def test_capacity_without_dropping_and_with_factor():
    cfg = C.MoEConfig(num_experts=8, expert_top_k=2, sequence_len=512, drop_tokens=0)
    assert cfg.EC == 512 * 2 and cfg.pEC == 1024
    cfg = C.MoEConfig(num_experts=8, expert_top_k=2, sequence_len=512, capacity_factor=2)
    assert cfg.EC == 64 * 2 * 2
    cfg = C.MoEConfig(num_experts=128, expert_top_k=2, sequence_len=8192, hidden_size=2048, intermediate_size=2048)
    assert cfg.PX == 128 and cfg.EC == 128 and cfg.TCM == 1


def test_expert_placement_is_contiguous_and_divisible():
    cfg = C.BASELINE_CONFIGS["C"]
    assert cfg.num_local_experts(8) == 1 and cfg.num_local_experts(4) == 2 and cfg.num_local_experts(1) == 8
    with pytest.raises(C.ConfigError):
        cfg.num_local_experts(3)


@pytest.mark.parametrize("bad", [
    {"hidden_size": 1000}, {"intermediate_size": 100}, {"sequence_len": 100}, {"torch_dtype": 7},
    {"drop_tokens": 2}, {"expert_top_k": 0}, {"hidden_act": 3}, {"capacity_factor": 0}, {"bogus_key": 1},
])
def test_schema_rejects(bad):
    raw = json.load(open(C.DEFAULT_CONFIG_PATH))
    raw.update(bad)
    with pytest.raises(C.ConfigError):
        C.from_dict(raw)


def test_missing_required_key_and_missing_file():
    raw = json.load(open(C.DEFAULT_CONFIG_PATH))
    del raw["num_experts"]
    with pytest.raises(C.ConfigError):
        C.from_dict(raw)
    with pytest.raises(FileNotFoundError):
        C.load_config("csrc/kleos_config.json")  # the reference's dangling default name


@pytest.mark.parametrize("bad", [{"torch_dtype": 1}, {"is_training": 2}, {"expert_top_k": 9, "num_experts": 16}])
def test_hot_path_constraints(bad):
    cfg = C.MoEConfig(**{**C.load_config().raw(), **bad})
    with pytest.raises(C.ConfigError):
        cfg.check_hot_path()
