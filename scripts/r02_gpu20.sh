#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r02_pytest_gpu.log 2>&1
tail -8 gpurun_out/r02_pytest_gpu.log
