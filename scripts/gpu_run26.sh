#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 6000 --csv --log-file gpurun_out/launches_probe.csv python scripts/probe.py --trees 2 --settle 40 > gpurun_out/probe_ncu.log 2>&1
tail -3 gpurun_out/probe_ncu.log
