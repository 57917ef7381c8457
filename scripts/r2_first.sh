#!/bin/bash
# round 2, first GPU call: parity of the restructured kernel, knock-out timings, trace, bench, cooperative launch under ncu
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== pytest gpu parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
echo "=== diag time default"; timeout 300 python scripts/diag.py --cfg B --stage time 2>&1 | tail -6
echo "=== diag time partner-arrive (dbg 16)"; FM_DBG_FLAGS=16 timeout 300 python scripts/diag.py --cfg B --stage time 2>&1 | tail -5
for dbg in 1 2 3 4 8; do
  echo "=== FFN knock-out dbg=$dbg (1 no loads, 2 no MMA, 4 no A, 8 no B)"
  FM_DBG_FLAGS=$dbg timeout 300 python scripts/diag.py --cfg B --stage time 2>&1 | grep -E "phase ffn|us/forward"
done
echo "=== trace"; timeout 300 python scripts/diag.py --cfg B --stage trace 2>&1 | tail -40
echo "=== bench"; timeout 600 python bench.py --steps 200 --warmup 32 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_first.json | cut -c1-600
echo "=== ncu launch list of smoke (cooperative default?)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 10 --csv --log-file gpurun_out/r2_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
tail -5 gpurun_out/r2_smoke_launches.csv
echo "=== coop outside ncu"; FM_COOP=1 timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
