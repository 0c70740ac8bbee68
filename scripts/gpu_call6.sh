#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout=200 -p no:cacheprovider --tb=short -k "tma_store or ffn_block or planes_gemm" 2>&1 | tail -5
for d in 0 1; do echo DIRECT=$d; DOST_GEMM_DIRECT=$d NBUF=3 timeout 200 python scripts/gemm_ffn_time.py DOST_GEMM_EPI16=0,1 2>&1 | tee -a gpurun_out/gemm_ffn_time_direct.txt; done
