#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 120 python scripts/probe.py --trees 8 --settle 40 2>&1 | tail -12
timeout 200 python scripts/longrun.py 300 2>&1 | tail -6
