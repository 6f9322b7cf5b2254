#!/bin/bash
# ncu --set full of selected kernels of one bench step.  usage: bash tools/gpu_ncu_kernel.sh <tag> <kernel regex> [count]
TAG=${1:-n}; O=gpurun_out/$TAG; mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$2" -c ${3:-2} \
   -o $O/full python bench.py --steps 1 --warmup 1 --cpu-sample 0 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O
