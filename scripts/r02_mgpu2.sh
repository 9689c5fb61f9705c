#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_baseline_shapes.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r02_pytest_gpu.log 2>&1
tail -8 gpurun_out/r02_pytest_gpu.log
python -m pytest tests/test_multi_gpu.py tests/test_host_sharding.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_mgpu.log 2>&1
tail -12 gpurun_out/r02_pytest_mgpu.log
