"""Expert-parallel world of 2 (and more when present) GPUs: dispatch / combine over NVLink peer memory inside the fused
kernel, checked per rank against the oracle.  Skipped on single-GPU boxes."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_expert_parallel_parity(world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-4000:] + "\n" + res.stderr[-4000:]
    for r in range(world):
        assert f"RANK {r} ALL OK" in res.stdout


def test_expert_parallel_on_torch_symmetric_memory():
    """The externally allocated symmetric slab (torch.distributed._symmetric_memory -> fm_symm_use_external +
    fm_symm_attach_ptrs) instead of the library's own cudaMalloc + CUDA IPC mapping; same parity bar, 2 GPUs."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    env["FM_SYMM"] = "torch"
    env["FM_MULTI_CASES"] = ('[{"num_experts":8,"expert_top_k":2,"sequence_len":512,"hidden_size":256,"intermediate_size":512},'
                             '{"num_experts":16,"expert_top_k":4,"sequence_len":256,"hidden_size":128,"intermediate_size":256,"hidden_act":1}]')
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stdout[-4000:] + "\n" + res.stderr[-4000:]
    assert "RANK 0 ALL OK" in res.stdout and "RANK 1 ALL OK" in res.stdout
