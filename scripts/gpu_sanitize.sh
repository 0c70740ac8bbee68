#!/bin/bash
# compute-sanitizer over the kernel parity tests incl. the fused attention kernel.  Summaries: gpurun_out/sanitizer_<tool>.log
mkdir -p gpurun_out
SEL='fused_attention or planes_gemm_forward_shapes or tma_store or ffn_block or cross_attention_tensor_core or segment_reduce or ln_planes or loss or csr_build or dense_attention_on_planes'
for TOOL in memcheck racecheck; do
  timeout ${SAN_TIMEOUT:-300} compute-sanitizer --tool $TOOL --print-limit 5 --error-exitcode 0 \
    python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout=280 -p no:cacheprovider --tb=line -k "$SEL" \
    > gpurun_out/sanitizer_$TOOL.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_$TOOL.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit " gpurun_out/sanitizer_$TOOL.log | tail -4
done
# initcheck, with the TMA-store epilogue on and off: are cp.async.bulk.tensor stores tracked as initialising writes?
for T in 1 0; do
  DOST_GEMM_TMA_EPI=$T timeout 200 compute-sanitizer --tool initcheck --print-limit 3 --error-exitcode 0 \
    python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout=180 -p no:cacheprovider --tb=line -k "ffn_block or ln_planes or segment_reduce or fused_attention_dense" \
    > gpurun_out/sanitizer_initcheck_tma$T.log 2>&1
  echo "TMA_EPI=$T: $(grep -c 'Uninitialized' gpurun_out/sanitizer_initcheck_tma$T.log) uninitialized-read reports; $(grep -E 'passed|failed' gpurun_out/sanitizer_initcheck_tma$T.log | tail -1)"
  grep "Device Frame" gpurun_out/sanitizer_initcheck_tma$T.log | sed 's/(.*//' | sort | uniq -c | head -5
done
