#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_graphed.py -q -m gpu --timeout=100 -p no:cacheprovider --tb=short -rf -x -k "tma_store or bucket_padding or batched_and_ragged or dense_attention_on_planes or tensor_core_formulation" > gpurun_out/pytest_tma.log 2>&1
rc=$?
echo "pytest exit $rc" >> gpurun_out/pytest_tma.log
tail -15 gpurun_out/pytest_tma.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests -q -m gpu --timeout=300 -p no:cacheprovider --tb=short -rf > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
B=512 TAG=r2_b512_b bash scripts/gpu_profile.sh > /dev/null
