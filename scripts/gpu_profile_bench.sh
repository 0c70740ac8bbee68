#!/bin/bash
# ncu launch list of ONE timed step of bench.py itself (the profiler window is opened by --ncu-window) + summary.
# The numbers bench.py prints under ncu are not bench values.
mkdir -p gpurun_out
TAG=${TAG:-cur}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_bench_${TAG}.csv python bench.py --steps 2 --warmup 3 --ncu-window --no-cpu-baseline --no-alt \
  > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_bench_${TAG}.csv > gpurun_out/launches_bench_${TAG}.summary.txt 2>&1
head -45 gpurun_out/launches_bench_${TAG}.summary.txt
