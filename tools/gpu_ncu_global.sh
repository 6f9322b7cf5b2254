#!/bin/bash
# ncu --set full of one launch of each named kernel of a global-mode compress.  usage: bash tools/gpu_ncu_global.sh <tag> <gap|divergent> <kernel> [<kernel> ...]
TAG=$1; SHAPE=$2; shift 2; O=gpurun_out/$TAG; mkdir -p $O
for K in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 3 -c 1 -o $O/$K python tools/one_global.py $SHAPE > $O/$K.log 2>&1; echo "$K rc=$?"
done
ls -la $O
