#!/bin/bash
mkdir -p gpurun_out
QR_KTRACE=20 timeout 300 python scripts/longrun_sharded.py 2 200 > gpurun_out/r02_ktrace15_n2.log 2>&1; tail -30 gpurun_out/r02_ktrace15_n2.log
QR_TRACE=1 timeout 300 python scripts/longrun_sharded.py 2 100 2>&1 | tail -45 > gpurun_out/r02_trace15_n2.log; tail -45 gpurun_out/r02_trace15_n2.log
