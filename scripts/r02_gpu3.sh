#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:'scan_kernel|route_kernel|leaf_node_kernel|hist_limb' --launch-skip 5600 -c 14 -o gpurun_out/r02_round_full -f python scripts/longrun.py 120 > gpurun_out/r02_ncu_full.log 2>&1
tail -3 gpurun_out/r02_ncu_full.log
ls -la gpurun_out/*.ncu-rep
