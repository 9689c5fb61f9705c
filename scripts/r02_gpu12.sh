#!/bin/bash
mkdir -p gpurun_out
( time python bench.py --config 4 --steps 10 --warmup 3 --e2e-trees 30 > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err ) 2>&1 | grep real
tail -3 gpurun_out/r02_bench_c4_n1.err; cut -c1-1500 gpurun_out/r02_bench_c4_n1.json
( time python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r02_bench_c3_n1.json 2> gpurun_out/r02_bench_c3_n1.err ) 2>&1 | grep real
tail -3 gpurun_out/r02_bench_c3_n1.err; cut -c1-1800 gpurun_out/r02_bench_c3_n1.json
