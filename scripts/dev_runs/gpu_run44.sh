#!/bin/bash
# 2-GPU: final check of the sharded path (squares prefetch in the fused scan) — parity in all modes + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_mgpu.log 2>&1
tail -6 gpurun_out/pytest_mgpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 50 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -1 gpurun_out/bench_n2.json | cut -c1-200; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n2.err | tail -4
echo "=== N=2, 1M docs (auto)"; timeout 200 python scripts/longrun_sharded.py 2 250 2>&1 | grep -E "trees|exchange"
