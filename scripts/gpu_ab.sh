#!/bin/bash
# same-box A/B of environment switches on the headline step: usage gpu_ab.sh "ENV1=a ENV2=b" "ENV1=c" ...  (each arg = one variant)
mkdir -p gpurun_out
: > gpurun_out/ab.txt
for rep in 1 2; do
for v in "$@"; do
  out=$(env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-alt 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])")
  echo "$v | $out" | tee -a gpurun_out/ab.txt
done
done
