#!/bin/bash
mkdir -p gpurun_out
QR_NO_PDL=1 ncu --set full --import-source on --clock-control none -k regex:'hist_limb' --launch-skip 1500 -c 6 -o gpurun_out/r02_fused_full -f python scripts/longrun.py 120 > gpurun_out/r02_ncu_full.log 2>&1
tail -2 gpurun_out/r02_ncu_full.log
ls -la gpurun_out/r02_fused_full.ncu-rep
