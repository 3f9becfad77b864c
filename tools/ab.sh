#!/bin/bash
# A/B of GEMM kernel modes on the full step: tools/ab.sh <batch> "ENV1=.. ENV2=.." "ENV..." ...
b=$1; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  echo "=== [$i] $envs" 
  env $envs python bench.py --steps 20 --warmup 5 --batch $b --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('ms/step', d['ms_per_step'], 'img/s', d['value'], 'gemm_ms', r['gemm_ms_per_step'], 'TF', r['achieved'])
    elif l: print(l[:300])
"
  env $envs python tools/gemm_breakdown.py betavaegan $b > gpurun_out/gemm_ab_$i.txt 2>&1
  i=$((i+1))
done
