#!/bin/bash
# 8-GPU validation: sharded parity at world 8 in the three peer-memory modes, bench at N=8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
QR_TEST_WORLD=8 timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x -k "LAMBDAMART and peer and not OBV" > gpurun_out/pytest_mgpu8.log 2>&1
tail -12 gpurun_out/pytest_mgpu8.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 30 --e2e-trees 300 --no-cpu-baseline > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -1 gpurun_out/bench_n8.json | cut -c1-400; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n8.err | tail -6
