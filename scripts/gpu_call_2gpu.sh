#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dp.py tests/test_gpu_guards.py tests/test_gpu_ops.py -q -m gpu --timeout=300 -p no:cacheprovider --tb=short -rf -k "two_gpu or non_current_device or dropout_on_tensor_cores or tma_store" > gpurun_out/pytest_2gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_2gpu.log
tail -12 gpurun_out/pytest_2gpu.log | cut -c1-300
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --extras-timeout 200 > gpurun_out/bench2.log 2> gpurun_out/bench2.err; echo "bench2 exit $?" >> gpurun_out/bench2.err
tail -3 gpurun_out/bench2.err | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --scaling strong > gpurun_out/bench2_strong.log 2> gpurun_out/bench2_strong.err; echo "exit $?" >> gpurun_out/bench2_strong.err
tail -2 gpurun_out/bench2_strong.err | cut -c1-300
