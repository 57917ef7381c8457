"""bench.py's output contract on the CPU-only legs: the reference arm prints one JSON line with the agreed keys, and the
multi-rank convention (only rank 0 works and prints) holds.  The GPU arm is exercised on the B200 box by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    env.update({"FM_BENCH_CPU_BUDGET_S": "2"})  # shrink the per-step sample so the test stays at seconds
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                           *args], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)


def test_reference_arm_json_line():
    res = _run()
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "moe_layer_fwd_tokens_per_s" and d["unit"] == "tokens/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "bf16" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "tokens" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # the same `config` object as the GPU arm prints (the driver compares the two arms' configs) ...
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench.config_dict("B", bench.BASELINE_CONFIGS["B"], 1)
    # ... and the product's native library is never mapped into the reference arm's process
    assert d["native_so_mapped"] == [] or all("flashmoe" not in s for s in d["native_so_mapped"]), d["native_so_mapped"]


def test_every_baseline_config_is_selectable():
    sys.path.insert(0, ROOT)
    import bench

    assert set(bench.BASELINE_CONFIGS) >= {"A", "B", "C", "D1k", "D4k", "D16k", "D64k", "E8", "E16", "E32", "E64", "E128"}
    for name, cfg in bench.BASELINE_CONFIGS.items():
        assert bench.config_dict(name, cfg, 8)["workload"].startswith("configs[")


def test_reference_arm_non_zero_ranks_exit_quietly():
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, args=("--gpus", "2"))
    assert res.returncode == 0 and not [l for l in res.stdout.splitlines() if l.startswith("{")]


def test_usable_cpus_respects_affinity():
    sys.path.insert(0, ROOT)
    import bench

    n = bench.usable_cpus()
    assert 1 <= n <= (os.cpu_count() or 1)
