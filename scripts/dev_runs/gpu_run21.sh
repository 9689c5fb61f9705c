#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 6000 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --settle 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_codes -s 1 -c 1 -o gpurun_out/r01_score_full python scripts/score_probe.py --n 1000000 --f 136 --trees 1000 > gpurun_out/score_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 1 -c 1 -o gpurun_out/r01_encode_full python scripts/score_probe.py --n 1000000 --f 136 --trees 1000 > gpurun_out/encode_ncu.log 2>&1
