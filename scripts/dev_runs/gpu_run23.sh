#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 7000 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 python scripts/probe.py --trees 8 --settle 40 2>&1 | tail -14
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --settle 30 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hist_limb -s 520 -c 14 -o gpurun_out/r01_hist_full python scripts/probe.py --trees 1 --settle 40 > gpurun_out/ncu_hist.log 2>&1
tail -2 gpurun_out/ncu_hist.log
ls -la gpurun_out
