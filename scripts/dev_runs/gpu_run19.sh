#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "scoring or score or host_cli" > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python scripts/score_probe.py --n 1000000 --f 700 --trees 5000 2>&1 | tail -2
timeout 600 python scripts/score_probe.py --n 1000000 --f 136 --trees 1000 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_codes -s 1 -c 1 -o gpurun_out/prof_score3 python scripts/score_probe.py --n 200000 --f 700 --trees 5000 > gpurun_out/score_ncu3.log 2>&1
