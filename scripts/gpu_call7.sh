#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout=60 -p no:cacheprovider --tb=short -x -k "fused_attention_dense" 2>&1 | tail -25 | cut -c1-220
timeout 180 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout=60 -p no:cacheprovider --tb=short -x -k "fused_attention_ragged" 2>&1 | tail -25 | cut -c1-220
