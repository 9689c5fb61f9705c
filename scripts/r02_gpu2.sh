#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_gpu.log 2>&1
tail -15 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 17000 -c 500 --csv --log-file gpurun_out/r02_launches_late.csv python scripts/longrun.py 200 > gpurun_out/r02_ncu_longrun.log 2>&1
tail -3 gpurun_out/r02_ncu_longrun.log
python scripts/summarize_launches.py gpurun_out/r02_launches_late.csv | head -30
