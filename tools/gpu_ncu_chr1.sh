#!/bin/bash
# ncu --set full of one kernel of the chr1-sized pair (second launch: warm).  usage: bash tools/gpu_ncu_chr1.sh <tag> <kernel regex> [launch-skip]
TAG=${1:-n}; O=gpurun_out/$TAG; mkdir -p $O
python tools/one_chr1.py 3 > $O/plain.log 2>&1; echo "plain rc=$?"; cat $O/plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" --launch-skip ${3:-1} -c 1 -o $O/full_$2 python tools/one_chr1.py 2 > $O/ncu_$2.log 2>&1; echo "ncu rc=$?"
ls -la $O
