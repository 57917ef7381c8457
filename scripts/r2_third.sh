#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== pytest gpu (all single-GPU files)"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "=== gantt full"; timeout 300 python scripts/trace_gantt.py --cfg B --label full 2>&1 | grep -v "p10" | head -40
timeout 300 python scripts/trace_gantt.py --cfg B --label full 2>&1 | grep -E "GEMM|per k-block|first landed->acc|per pair"
for ca in 4 12 14; do echo "=== claim ahead $ca"; FM_CLAIM_AHEAD_KB=$ca timeout 300 python scripts/trace_gantt.py --cfg B --label ca$ca 2>&1 | grep -E "^===|per pair"; done
echo "=== dbg=3"; FM_DBG_FLAGS=3 timeout 300 python scripts/trace_gantt.py --cfg B --label dbg3 2>&1 | grep -E "^===|per pair|per k-block"
echo "=== bench B"; timeout 600 python bench.py --steps 200 --warmup 32 2>&1 | tail -1 | tee gpurun_out/r2_bench_b.json | cut -c1-400
echo "=== reference arm"; FM_BENCH_CPU_BUDGET_S=30 timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-900
