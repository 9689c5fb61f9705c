#!/bin/bash
mkdir -p gpurun_out
QR_KTRACE=20 timeout 300 python scripts/longrun.py 200 > gpurun_out/r02_ktrace14.log 2>&1; tail -30 gpurun_out/r02_ktrace14.log
QR_KTRACE=3 timeout 300 python scripts/longrun.py 100 > gpurun_out/r02_ktrace14b.log 2>&1; tail -12 gpurun_out/r02_ktrace14b.log
