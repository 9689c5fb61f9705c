#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python scripts/probe.py --trees 10 > gpurun_out/probe_fast.log 2>&1; tail -14 gpurun_out/probe_fast.log
timeout 600 python scripts/probe.py --trees 3 --mode 1 > gpurun_out/probe_exact.log 2>&1; tail -12 gpurun_out/probe_exact.log
