#!/bin/bash
# One profiling pass for profiles/: launch list of the bench command, then full-set captures of the hot kernels.
set -x
R=${ROUND:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --headline-only > gpurun_out/${R}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vbq_bisect_tma_kernel -s 3 -c 1 -o gpurun_out/${R}_quantize \
    python bench.py --steps 3 --warmup 3 --no-cpu --headline-only > gpurun_out/${R}_quantize.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vbq_bisect_tma_kernel -s 3 -c 1 -o gpurun_out/${R}_corrected \
    python scripts/run_both_once.py > gpurun_out/${R}_corrected.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vbq_bisect_sweep -s 2 -c 1 -o gpurun_out/${R}_sweep \
    python scripts/bench_configs.py sweep > gpurun_out/${R}_sweep.log 2>&1
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
tail -c 600 gpurun_out/${R}_bench.json
