#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_linesearch.py tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_ls.log 2>&1; tail -5 gpurun_out/r02_pytest_ls.log
timeout 600 python scripts/linesearch_probe.py 200000 50
