#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout=300 -p no:cacheprovider --tb=short -rf > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
# racecheck: every hazard (print limit high), our kernels are filtered here afterwards
SEL='ffn_block or cross_attention_tensor_core or segment_reduce or ln_planes or csr_build or cross_attention_matches_padded or tma_store'
timeout 400 compute-sanitizer --tool racecheck --print-limit 100000 --error-exitcode 0 python -m pytest tests/test_gpu_ops.py -q -m gpu -x --timeout=380 -p no:cacheprovider --tb=line -k "$SEL" > gpurun_out/sanitizer_racecheck_all.log 2>&1
grep -E "Race reported|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck_all.log | grep -v "at::native" | tail -8 | cut -c1-300
# initcheck of the FFN block with the TMA-store epilogue on and off (are bulk-tensor stores tracked as initialising writes?)
for T in 1 0; do
DOST_GEMM_TMA_EPI=$T timeout 300 compute-sanitizer --tool initcheck --print-limit 3 --error-exitcode 0 python -m pytest tests/test_gpu_ops.py -q -m gpu -x --timeout=280 -p no:cacheprovider --tb=line -k "ffn_block or ln_planes or segment_reduce" > gpurun_out/sanitizer_initcheck_tma$T.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Uninitialized" gpurun_out/sanitizer_initcheck_tma$T.log | sort | uniq -c | tail -4
done
