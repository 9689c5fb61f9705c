#!/bin/bash
# 2-GPU: growth-round timeline of the sharded path, and how the exchange pays off on a larger dataset
mkdir -p gpurun_out
echo "=== N=2, 1M docs, traced"; QR_TRACE=1 QR_TRACE_ROUNDS=1 timeout 200 python scripts/longrun_sharded.py 2 201 2>&1 | grep -E "trace|trees|exchange" | tail -42
echo "=== N=1, 1M docs, traced"; QR_TRACE=1 QR_TRACE_ROUNDS=1 timeout 200 python scripts/longrun_sharded.py 1 201 2>&1 | grep -E "trace|trees|exchange" | tail -38
echo "=== N=2, 1M docs"; timeout 200 python scripts/longrun_sharded.py 2 300 2>&1 | grep -E "trees|exchange"
echo "=== N=1, 4M docs"; timeout 300 python scripts/longrun_sharded.py 1 150 4000000 2>&1 | grep -E "trees|exchange"
echo "=== N=2, 4M docs"; timeout 300 python scripts/longrun_sharded.py 2 150 4000000 2>&1 | grep -E "trees|exchange"
