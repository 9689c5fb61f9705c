#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "edge_case" > gpurun_out/pytest_edge.log 2>&1
tail -30 gpurun_out/pytest_edge.log | cut -c1-220
