#!/bin/bash
# `ncu --set full` of the dominant GEMM with the step's own epilogues (one launch each), DRAM traffic per launch -> JSON.
# Output: gpurun_out/r2_ncu_<name>.ncu-rep (+ .raw.csv); scripts/ncu_extract.py turns them into profiles/ summaries here.
mkdir -p gpurun_out
for NAME in ${NAMES:-fc1_fwd fc2_fwd fc2_dA}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf_kernel -s 2 -c 1 -f \
    -o gpurun_out/r2_ncu_${NAME} python scripts/gemm_ffn_one.py ${NAME} bf16x3 3 > gpurun_out/r2_ncu_${NAME}.log 2>&1
  tail -2 gpurun_out/r2_ncu_${NAME}.log
done
