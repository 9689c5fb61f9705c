#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r02_pytest_mgpu.log 2>&1
tail -8 gpurun_out/r02_pytest_mgpu.log
timeout 300 python scripts/longrun_sharded.py 2 200 2>&1 | tail -5
QR_KTRACE=20 timeout 300 python scripts/longrun_sharded.py 2 200 2>&1 | grep -A3 "tree 149" | tail -5
