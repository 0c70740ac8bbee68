#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout=120 -p no:cacheprovider --tb=short -rf -k "tma_store or planes_gemm or ffn_block or edge_block" > gpurun_out/pytest_tma.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tma.log
tail -15 gpurun_out/pytest_tma.log
timeout 900 python -m pytest tests -q -m gpu --timeout=600 -p no:cacheprovider --tb=short -rf > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
