#!/bin/bash
# Runs on the GPU box: alternating bench runs of library variants (same box, same clocks).  usage: gpu_ab.sh TAG variant...
# ("main" = slimt_b200/libslimt_b200.so, anything else = slimt_b200/libslimt_b200_<variant>.so)
TAG=$1; shift
for rep in 1 2; do
  for v in "$@"; do
    if [ "$v" = "main" ]; then unset SLIMT_B200_LIB; else export SLIMT_B200_LIB=$PWD/slimt_b200/libslimt_b200_$v.so; fi
    timeout 600 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/${TAG}_${v}_${rep}.json 2> gpurun_out/${TAG}_${v}_${rep}.err
    echo "== $v rep $rep"; python tools/bench_summary.py gpurun_out/${TAG}_${v}_${rep}.json | head -7
  done
done
