#!/bin/bash
# Runs on a multi-GPU box: weak scaling (one process per GPU, torchrun) of the non-headline configs and strong scaling
# of ONE request served by one process (slimt_b200_translate_multi).   tools/gpu_multi_r2.sh TAG NGPUS [quick]
TAG=${1:-r2m}; N=${2:-8}; QUICK=${3:-}
show() { python - "$1" <<'EOF'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "n_gpus", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), "cpu", d.get("cpu_baseline",{}).get("value"), d.get("mode",""))
except Exception as e: print(sys.argv[1], "failed", e)
EOF
}
tr() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/${TAG}_n${N}_${name}.json 2> gpurun_out/${TAG}_n${N}_${name}.err; show gpurun_out/${TAG}_n${N}_${name}.json; }
sp() { n=$1; sents=$2; timeout 900 python bench.py --single-process --gpus $n --workload mixed --sentences $sents --steps 1 --warmup 3 > gpurun_out/${TAG}_sp_n${n}_s${sents}.json 2> gpurun_out/${TAG}_sp_n${n}_s${sents}.err; show gpurun_out/${TAG}_sp_n${n}_s${sents}.json; }
if [ -n "$QUICK" ]; then
  tr mixed --workload mixed --steps 1 --warmup 2 --math exact --no-cpu-baseline
  sp $N 65536
  exit 0
fi
tr base_shortlist --workload base_shortlist --steps 10 --warmup 3 --math exact
tr mixed --workload mixed --steps 2 --warmup 2 --math exact
tr tiny_full --workload tiny_full --steps 30 --warmup 3 --math exact --no-cpu-baseline
sp $N 1000000
sp 4 1000000
for n in 8 4 2 1; do sp $n 262144; done
