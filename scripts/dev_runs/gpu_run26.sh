#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 27000 --launch-count 1200 --csv --log-file gpurun_out/launches_late.csv python scripts/probe.py --trees 1 --settle 300 > gpurun_out/probe_ncu.log 2>&1
tail -3 gpurun_out/probe_ncu.log
