#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r02_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu.log
timeout 300 python scripts/longrun.py 250 2>&1 | tail -5
QR_ROW_COPY=0 timeout 300 python scripts/longrun.py 250 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err; cut -c1-300 gpurun_out/r02_bench_n1.json
