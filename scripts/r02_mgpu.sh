#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_multi_gpu.py tests/test_host_sharding.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r02_pytest_mgpu.log 2>&1
tail -15 gpurun_out/r02_pytest_mgpu.log
timeout 300 python scripts/longrun_sharded.py 2 150 2>&1 | tail -5
