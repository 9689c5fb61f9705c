#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_linesearch.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_ls.log 2>&1; tail -30 gpurun_out/r02_pytest_ls.log
