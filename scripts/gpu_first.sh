#!/bin/bash
# first GPU contact: parity tests (all, no -x), smoke, phase probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python scripts/probe.py --trees 5 > gpurun_out/probe_fast.log 2>&1; tail -15 gpurun_out/probe_fast.log
