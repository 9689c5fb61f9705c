#!/bin/bash
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 5 --e2e-trees 400 > gpurun_out/r02_bench_c2_n${N}_b.json 2> gpurun_out/r02_bench_c2_n${N}_b.err
tail -2 gpurun_out/r02_bench_c2_n${N}_b.err | cut -c1-300; grep '^{' gpurun_out/r02_bench_c2_n${N}_b.json | cut -c1-300
