#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r02_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu.log
timeout 300 python scripts/longrun.py 300 > gpurun_out/r02_longrun.log 2>&1; tail -6 gpurun_out/r02_longrun.log
echo "--- no PDL"; QR_NO_PDL=1 timeout 300 python scripts/longrun.py 300 2>&1 | tail -2
QR_KTRACE=20 timeout 300 python scripts/longrun.py 250 2>&1 | grep ktrace | tail -6
