#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "scoring or score or host_cli" > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python scripts/score_sweep.py --n 1000000 --f 136 --trees 1000 --shapes auto,1:1,1:2,2:1,2:2,4:1,4:2,4:4,1:1:512,1:1:256,2:1:256,2:1:128,4:1:128 2>&1 | tail -14
timeout 900 python scripts/score_sweep.py --n 1000000 --f 700 --trees 5000 --shapes auto,2:1,2:2,4:1,4:2,4:4,1:1 2>&1 | tail -8
