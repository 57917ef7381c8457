#!/bin/bash
# single GPU: parity suite, Gantt prints with the finer router stamps (B, E32 = D4k shape, E128, D16k), bench B
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== pytest gpu parity"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_large_shapes.py -x -q -m gpu 2>&1 | tail -6
F='^===|per pair|gate_done|disp_prefix|dispatch_end|kernel_end|gate_topk|gate_gemv|disp_rows|zero_issued|topk_warp0|barrier|first_tma'
for c in B D4k E128 D16k; do
  echo "=== gantt $c"; timeout 300 python scripts/trace_gantt.py --cfg $c --label $c 2>&1 | grep -E "$F"
done
echo "=== bench B"; timeout 600 python bench.py --steps 200 --warmup 32 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_b6.json | cut -c1-300
