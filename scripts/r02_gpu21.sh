#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_baseline_shapes.py -m gpu -q --tb=short -p no:cacheprovider -x -k "reference or ordered or smoke or oblivious or REFERENCE or config" > gpurun_out/r02_pytest_ref.log 2>&1
tail -15 gpurun_out/r02_pytest_ref.log
timeout 600 python scripts/refmode_probe.py 8 2>&1 | tail -12
QR_EXACT_WALK_MIN=1 timeout 600 python scripts/refmode_probe.py 6 2>&1 | tail -4
QR_EXACT_WALK_MIN=250000 timeout 600 python scripts/refmode_probe.py 6 2>&1 | tail -4
