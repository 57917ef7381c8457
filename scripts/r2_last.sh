#!/bin/bash
# single GPU: parity suite incl. the multi-sub-chunk tensor-core router cases; the 64k-token sweep point with its parity check
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== pytest gpu parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -6
echo "=== bench D64k (1 GPU), parity on 128 tokens"
timeout 300 python bench.py --config D64k --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --parity-tokens 128 2>&1 | tail -1 | tee gpurun_out/r2_bench_d64k_n1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); pc=d['parity_check']; print(d['ms_per_step'], {k:pc[k] for k in pc if k not in ('rule','inputs')})"
