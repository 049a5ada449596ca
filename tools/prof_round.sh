#!/bin/bash
# Runs on the GPU box (gpurun): launch list of one bench pass + ncu --set full captures of the main kernels.
# Outputs land in gpurun_out/; tools/prof_collect.py (run on the CPU box) turns them into profiles/.
set -x
TAG=${1:-r1}
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
# only the timed pass is listed: bench.py brackets it with cudaProfilerStart/Stop
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH --profiler-range > gpurun_out/${TAG}_launches.log 2>&1
for spec in "cross_attention_rc:100:1" "out_argmax:20:1" "rows_ffn_kernel:0:1" "rows_ffn_kernel:8:1" "dec_ssru_kernel:8:1" "enc_attention_warp:2:1"; do
  IFS=: read pat skip cnt <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -o gpurun_out/${TAG}_${pat}_s${skip} $BENCH > gpurun_out/${TAG}_${pat}_s${skip}.log 2>&1
done
ls -la gpurun_out | tail -20
