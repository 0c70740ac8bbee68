#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --timeout=600 -p no:cacheprovider --tb=short -rf > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python scripts/graph_probe.py 8 64 512 > gpurun_out/graph_probe.log 2>&1; echo "exit $?" >> gpurun_out/graph_probe.log
tail -8 gpurun_out/graph_probe.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench2.log 2> gpurun_out/bench2.err; echo "bench2 exit $?" >> gpurun_out/bench2.err
tail -3 gpurun_out/bench2.err
