#!/bin/bash
# Runs on the GPU box: the round's closing evidence -- GPU parity suite, smoke, the default bench line, the reference
# arm, and the ncu launch list of one timed pass.  usage: gpu_final.sh TAG
TAG=${1:-r3}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; python tools/bench_summary.py gpurun_out/${TAG}_bench_n1.json | head -9
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 400 gpurun_out/${TAG}_bench_reference.json
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --math exact"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH --profiler-range > gpurun_out/${TAG}_launches.log 2>&1
tail -2 gpurun_out/${TAG}_launches.log   # tools/prof_collect.py TAG (CPU side) turns the launch list into profiles/TAG_launches_summary.txt
