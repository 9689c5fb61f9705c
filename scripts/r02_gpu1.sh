#!/bin/bash
# round 2, first GPU contact of the route -> histogram -> scan pipeline: parity suite, smoke, long-run timing + trace
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r02_pytest_gpu.log 2>&1
tail -15 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
timeout 300 python scripts/longrun.py 400 > gpurun_out/r02_longrun.log 2>&1; tail -10 gpurun_out/r02_longrun.log
QR_TRACE=1 QR_TRACE_ROUNDS=1 timeout 300 python scripts/longrun.py 260 2> gpurun_out/r02_trace.log | tail -3
grep "rounds:" gpurun_out/r02_trace.log | tail -5
