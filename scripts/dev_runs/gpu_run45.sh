#!/bin/bash
# 1-GPU: round-1 final state — parity suite, smoke, bench line, ncu launch list, ncu --set full of the histogram kernel
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -1 gpurun_out/bench_n1.json | cut -c1-200; tail -2 gpurun_out/bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --settle 30 --e2e-trees 2 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
wc -l gpurun_out/r01_launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hist_limb -s 500 -c 14 -o gpurun_out/r01_hist_full -f python scripts/probe.py --trees 1 --settle 40 > gpurun_out/ncu_hist.log 2>&1
tail -2 gpurun_out/ncu_hist.log
