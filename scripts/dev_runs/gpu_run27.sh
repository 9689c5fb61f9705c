#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 9000 gpurun_out/bench_n1.json | cut -c1-2600; tail -3 gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --settle 30 --e2e-trees 2 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hist_limb -s 500 -c 14 -o gpurun_out/r01_hist_full python scripts/probe.py --trees 1 --settle 40 > gpurun_out/ncu_hist.log 2>&1
tail -2 gpurun_out/ncu_hist.log
