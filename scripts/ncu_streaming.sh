#!/bin/bash
# `ncu --set full` of the HBM-bound kernels of one training step (LayerNorm forward / backward on the token and the edge
# streams, the CSR segmented reductions): dram bytes per launch against the algorithmic bytes.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"ln_fwd_kernel|ln_bwd_kernel|segment_reduce_vec_kernel|ds_from_planes" -c 40 -f -o gpurun_out/r2_ncu_streaming \
  python scripts/profile_step.py > gpurun_out/r2_ncu_streaming.log 2>&1
tail -2 gpurun_out/r2_ncu_streaming.log
