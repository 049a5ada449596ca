set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 2409 -c 2500 --csv --log-file gpurun_out/r1_base_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1_base_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cross_attention -s 100 -c 2 -o gpurun_out/r1_base_cross_attention python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1_base_ca.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_i8_kernel -s 400 -c 12 -o gpurun_out/r1_base_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1_base_gemm.log 2>&1
ls -la gpurun_out
