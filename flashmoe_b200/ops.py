"""User-facing entry points, same signatures as the reference's flashmoe/ops.py:18-71."""
from __future__ import annotations

from typing import Dict, Optional

from . import config as _config
from .launcher import launch_workers


def run_moe(n_processes: int = 1, processes_per_node: Optional[int] = None, hostfile: Optional[str] = None,
            config_path: str = str(_config.DEFAULT_CONFIG_PATH)):
    """Run one MoE layer forward on synthetic tensors sized by the compiled configuration.

    n_processes        one process per GPU (1 = single GPU)
    processes_per_node defaults to n_processes (single node)
    hostfile           accepted for signature compatibility; multi-node launches are out of scope for this build
    config_path        must describe the configuration the library was built with
                       (default csrc/flashmoe_config.json; the reference's default names a file that does not exist)
    Returns the launcher's CompletedProcess (the reference returns None and only prints).
    """
    if processes_per_node is None:
        processes_per_node = n_processes
    return launch_workers(config_path=config_path, n_processes=n_processes, processes_per_node=processes_per_node,
                          hostfile=hostfile)


def get_compiled_config() -> Dict[str, int]:
    """Compile-time dimensions: keys S, H, E, P, PX, Element_size (what `_C.get_compiled_config` really returns in the
    reference, python_bindings.cu:170-179; its docstring's `element_size_bytes` key never existed)."""
    from . import _C  # raises if the native library is missing

    return _C.get_compiled_config()
