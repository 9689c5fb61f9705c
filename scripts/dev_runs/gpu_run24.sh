#!/bin/bash
mkdir -p gpurun_out
timeout 300 ./scripts/hist_mb2 1000000 0.3 2>&1 | tee gpurun_out/hist_mb2.log
echo "---- 5% gather"
timeout 300 ./scripts/hist_mb2 1000000 0.05 2>&1 | grep gather | tee -a gpurun_out/hist_mb2.log
