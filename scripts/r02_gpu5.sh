#!/bin/bash
mkdir -p gpurun_out
QR_KTRACE=20 timeout 300 python scripts/longrun.py 250 > gpurun_out/r02_ktrace.log 2>&1; grep ktrace gpurun_out/r02_ktrace.log | tail -12
QR_KTRACE=5 timeout 300 python scripts/longrun.py 250 2>&1 | grep ktrace | tail -4
QR_KTRACE=0 timeout 300 python scripts/longrun.py 250 2>&1 | grep ktrace | tail -4
