#!/bin/bash
# ncu evidence for round 2 (one GPU): launch list of the bench command + one full capture of the fused kernel
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
BENCH="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --parity-tokens 0"
echo "=== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5 -c 20 --csv --log-file gpurun_out/r02_launches_configB.csv $BENCH > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r02_launches_configB.csv | cut -c1-300
echo "=== full capture"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:fm_moe_forward -s 8 -c 2 -f -o gpurun_out/r02_prof $BENCH > gpurun_out/r02_bench_under_ncu_full.log 2>&1
ls -la gpurun_out/r02_prof.ncu-rep
