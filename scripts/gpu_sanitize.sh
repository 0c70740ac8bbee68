#!/bin/bash
# compute-sanitizer over the kernel parity tests.  Summaries land in gpurun_out/sanitizer_<tool>.log;
# scripts/sanitizer_summary.py condenses them into profiles/r2_sanitizer.summary.txt here.
mkdir -p gpurun_out
SEL='planes_gemm_forward_shapes or tma_store or ffn_block or cross_attention_tensor_core or segment_reduce or ln_planes or loss or csr_build or dense_attention_on_planes or cross_attention_matches_padded'
for TOOL in memcheck racecheck initcheck; do
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $TOOL --print-limit 5 --error-exitcode 0 \
    python -m pytest tests/test_gpu_ops.py -q -m gpu -x --timeout=400 -p no:cacheprovider --tb=line -k "$SEL" \
    > gpurun_out/sanitizer_$TOOL.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_$TOOL.log
  grep -E "ERROR SUMMARY|passed|failed|exit " gpurun_out/sanitizer_$TOOL.log | tail -4
done
