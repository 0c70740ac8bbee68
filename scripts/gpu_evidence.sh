#!/bin/bash
# Evidence run (1 GPU): ncu --set full of the FFN GEMMs with the step's epilogues, the launch list of one timed step of
# bench.py itself (--ncu-window), compute-sanitizer over the kernel tests.
mkdir -p gpurun_out
bash scripts/ncu_gemm_traffic.sh
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --ncu-window --no-extras --no-cpu-baseline --no-alt \
  > gpurun_out/r2_bench_under_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2_launches_bench.csv > gpurun_out/r2_launches_bench.summary.txt 2>&1
head -12 gpurun_out/r2_launches_bench.summary.txt
SAN_TIMEOUT=240 bash scripts/gpu_sanitize.sh
