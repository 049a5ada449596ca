#!/bin/bash
# Runs on the GPU box: one bench line per non-headline BASELINE.json GPU config.  usage: gpu_workloads.sh TAG
TAG=${1:-wl}
for w in tiny_full base_shortlist mixed; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/${TAG}_${w}.json 2> gpurun_out/${TAG}_${w}.err
  echo "== $w"; tail -2 gpurun_out/${TAG}_${w}.err; python tools/bench_summary.py gpurun_out/${TAG}_${w}.json 2>/dev/null | head -9
done
