#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/probe.py --trees 10 --settle 40 > gpurun_out/probe_fast.log 2>&1; tail -18 gpurun_out/probe_fast.log
