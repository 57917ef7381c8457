#!/bin/bash
# the driver's round-end sequence on one GPU with the final build: pytest -m gpu, smoke(), bench.py (defaults)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench.py --steps 20 --warmup 5"; timeout 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/r2_bench_b_n1_final.json | cut -c1-2500
