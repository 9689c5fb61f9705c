#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --e2e-trees 100 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 3000 gpurun_out/bench_n2.json | cut -c1-1800; tail -5 gpurun_out/bench_n2.err
