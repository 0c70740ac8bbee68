#!/bin/bash
# ncu launch list of one training step (cudaProfilerStart/Stop window) + summary
mkdir -p gpurun_out
TAG=${TAG:-cur}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_${TAG}.csv python scripts/profile_step.py > gpurun_out/profile_step_${TAG}.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_${TAG}.csv > gpurun_out/launches_${TAG}.summary.txt 2>&1
head -45 gpurun_out/launches_${TAG}.summary.txt
