#!/bin/bash
# Runs on the GPU box: the non-headline BASELINE.json configs through bench.py (one GPU), both arithmetic modes.
TAG=${1:-r2d}
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err; tail -c 200 gpurun_out/${TAG}_bench_${name}.err; python - <<EOF
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
    print("${name}", round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "fast", round(d.get("fast",{}).get("value",0)), d.get("fast_vs_exact",{}).get("step_tokens_equal"), "cpu", d.get("cpu_baseline",{}).get("value"))
except Exception as e: print("${name} failed", e)
EOF
}
run mixed --workload mixed --steps 3 --warmup 3
run base_shortlist --workload base_shortlist --steps 10 --warmup 3
run tiny_full --workload tiny_full --steps 30 --warmup 3
run tiny_b64 --workload tiny_b64 --steps 100 --warmup 5
run tiny_len64 --workload tiny_len64 --steps 20 --warmup 3
run sp_mixed_n1 --single-process --gpus 1 --workload mixed --sentences 65536 --steps 1 --warmup 3
