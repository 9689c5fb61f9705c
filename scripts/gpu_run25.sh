#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
QR_INIT_TIMING=1 timeout 120 python scripts/probe.py --trees 8 --settle 40 2>&1 | tail -28
