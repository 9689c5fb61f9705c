#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_ref_launches.csv python scripts/refmode_probe.py 3 > gpurun_out/r02_ref_ncu.log 2>&1
tail -3 gpurun_out/r02_ref_ncu.log
python scripts/summarize_launches.py gpurun_out/r02_ref_launches.csv | head -40
