#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_gpu.log 2>&1
tail -12 gpurun_out/r02_pytest_gpu.log
timeout 300 python scripts/longrun.py 300 > gpurun_out/r02_longrun.log 2>&1; tail -3 gpurun_out/r02_longrun.log
bash scripts/r02_ncu_all.sh
