#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x -k "auto or unsliced" > gpurun_out/r02_pytest_mgpu.log 2>&1
tail -4 gpurun_out/r02_pytest_mgpu.log
QR_KTRACE=20 timeout 300 python scripts/longrun_sharded.py 2 200 2>&1 | grep -v "tree 99\|tree 49" | tail -12
