#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host_cli.py -m gpu -q --tb=short -p no:cacheprovider -k dart > gpurun_out/pytest_host.log 2>&1
tail -40 gpurun_out/pytest_host.log
