#!/bin/bash
# Runs on the GPU box: ncu --set full captures of selected kernels in one arithmetic mode.
#   tools/prof_r2.sh TAG MATH "pat:skip:cnt ..."
TAG=${1:-r2}; MATH=${2:-exact}; SPECS=${3:-"out_argmax:20:1"}
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --math $MATH"
for spec in $SPECS; do
  IFS=: read pat skip cnt <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -o gpurun_out/${TAG}_${MATH}_${pat}_s${skip} $BENCH > gpurun_out/${TAG}_${MATH}_${pat}_s${skip}.log 2>&1
done
ls -la gpurun_out | tail -8
