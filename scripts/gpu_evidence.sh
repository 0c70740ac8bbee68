#!/bin/bash
# Round-2 evidence (1 GPU): ncu --set full of the fused attention kernel (self and cross) and of the FFN GEMMs with the
# step's epilogues, the launch list of one timed step of bench.py itself, cuobjdump mnemonic counts of the shipped library.
mkdir -p gpurun_out
for NAME in self cross; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 1 -c 1 -f \
    -o gpurun_out/r2_ncu_attn_${NAME} python scripts/attn_one.py ${NAME} 3 > gpurun_out/r2_ncu_attn_${NAME}.log 2>&1
  tail -1 gpurun_out/r2_ncu_attn_${NAME}.log
done
NAMES="fc1_fwd fc2_fwd fc2_dA" bash scripts/ncu_gemm_traffic.sh
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r2_launches_bench_v2.csv python bench.py --steps 3 --warmup 3 --ncu-window --no-extras --no-cpu-baseline --no-alt \
  > gpurun_out/r2_bench_under_ncu_v2.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2_launches_bench_v2.csv > gpurun_out/r2_launches_bench_v2.summary.txt 2>&1
head -14 gpurun_out/r2_launches_bench_v2.summary.txt
