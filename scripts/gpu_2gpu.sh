#!/bin/bash
# 2 GPUs: every GPU test (the data-parallel parity tests run instead of skipping), then the 2-GPU bench lines (weak + strong)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout=300 -p no:cacheprovider --tb=short -rf > gpurun_out/pytest_2gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_2gpu.log
tail -6 gpurun_out/pytest_2gpu.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-extras --no-alt --no-cpu-baseline > gpurun_out/bench2.log 2> gpurun_out/bench2.err; echo "bench2 exit $?" >> gpurun_out/bench2.err
tail -2 gpurun_out/bench2.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/bench2.log').read().strip().splitlines()[-1]); print('weak2', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --scaling strong --no-extras --no-alt --no-cpu-baseline > gpurun_out/bench2_strong.log 2> gpurun_out/bench2_strong.err; echo "exit $?" >> gpurun_out/bench2_strong.err
python -c "
import json
d=json.loads(open('gpurun_out/bench2_strong.log').read().strip().splitlines()[-1]); print('strong2', d['value'], d['ms_per_step'], d['e2e']['value'])"
