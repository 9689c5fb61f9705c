#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:finalize_kernel --launch-skip 7000 -c 3 -o gpurun_out/fin_full python scripts/probe.py --trees 1 --settle 300 > gpurun_out/ncu_fin.log 2>&1
tail -2 gpurun_out/ncu_fin.log
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:partition_onepass --launch-skip 7000 -c 3 -o gpurun_out/part_full python scripts/probe.py --trees 1 --settle 300 > gpurun_out/ncu_part.log 2>&1
tail -2 gpurun_out/ncu_part.log
