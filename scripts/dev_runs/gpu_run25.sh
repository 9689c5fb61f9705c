#!/bin/bash
mkdir -p gpurun_out
timeout 100 python scripts/longrun.py 150 2>&1 | tail -3
timeout 120 python scripts/probe.py --trees 8 --settle 40 2>&1 | grep -E "unprofiled|warm"
