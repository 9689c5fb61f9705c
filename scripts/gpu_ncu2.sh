#!/bin/bash
mkdir -p gpurun_out
# steady-state tree: skip the launches of 40 settle trees (~45 launches each); capture a few hist launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hist_limb -s 520 -c 6 -o gpurun_out/prof_hist2 python scripts/probe.py --trees 1 --settle 40 > gpurun_out/ncu_hist2.log 2>&1
tail -2 gpurun_out/ncu_hist2.log
