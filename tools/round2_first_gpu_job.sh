#!/bin/bash
# First gpurun call of the next round: run what round 1 could only verify on CPU, then the usual evidence.
#   gpurun --timeout 2400 -- 'bash tools/round2_first_gpu_job.sh'
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
BG_RUN_UNVERIFIED=1 python -m pytest tests/test_zz_reference_golden_gpu.py -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_unverified.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --quick --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1 || true
tail -5 gpurun_out/pytest_gpu.log gpurun_out/pytest_unverified.log gpurun_out/bench_1gpu.json
# one full-set capture of the dominant kernel (dram traffic for roofline.traffic; tensor-pipe activity)
ncu --set full --clock-control none --import-source on -k regex:tc_fwd_kernel -s 60 -c 3 -o gpurun_out/prof_tc_fwd \
    python bench.py --steps 1 --warmup 1 --no-graph --quick --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1 || true
ncu -i gpurun_out/prof_tc_fwd.ncu-rep --page raw --csv 2>/dev/null | head -c 200000 > gpurun_out/prof_tc_fwd_raw.csv || true
