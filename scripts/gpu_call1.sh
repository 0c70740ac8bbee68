#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --timeout=600 -p no:cacheprovider --tb=short -rf -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python scripts/graph_probe.py 8 64 512 > gpurun_out/graph_probe.log 2>&1; echo "exit $?" >> gpurun_out/graph_probe.log
tail -8 gpurun_out/graph_probe.log
