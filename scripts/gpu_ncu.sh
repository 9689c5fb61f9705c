#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hist_limb -s 6 -c 4 -o gpurun_out/prof_hist python scripts/probe.py --trees 1 > gpurun_out/ncu_hist.log 2>&1
tail -3 gpurun_out/ncu_hist.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
wc -l gpurun_out/launches.csv
ls -la gpurun_out/
