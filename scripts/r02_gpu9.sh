#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r02_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu.log
timeout 300 python scripts/longrun.py 300 > gpurun_out/r02_longrun.log 2>&1; tail -6 gpurun_out/r02_longrun.log
QR_KTRACE=20 timeout 300 python scripts/longrun.py 250 2>&1 | grep ktrace | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 13000 -c 400 --csv --log-file gpurun_out/r02_launches_late.csv python scripts/longrun.py 200 > gpurun_out/r02_ncu_longrun.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_launches_late.csv | head -24
