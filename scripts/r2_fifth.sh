#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== pytest gpu parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
F='^===|per k-block|per pair|gate_done|disp_prefix|dispatch_end|ffn_end|kernel_end|barrier|gate_topk|gate_gemv|disp_rows'
echo "=== gantt"; timeout 300 python scripts/trace_gantt.py --cfg B --label default 2>&1 | grep -E "$F"
echo "=== bench B"; timeout 600 python bench.py --steps 200 --warmup 32 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_b3.json | cut -c1-300
