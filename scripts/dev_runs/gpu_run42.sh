#!/bin/bash
# 2-GPU: sharded parity in all exchange modes, bench at N=2 per mode, round timeline of the fused exchange
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_mgpu.log 2>&1
tail -15 gpurun_out/pytest_mgpu.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "two_million" > gpurun_out/pytest_big.log 2>&1
tail -5 gpurun_out/pytest_big.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 50 --e2e-trees 200 --no-cpu-baseline > gpurun_out/bench_n2_$name.json 2> gpurun_out/bench_n2_$name.err
  echo "== $name"; tail -1 gpurun_out/bench_n2_$name.json | cut -c1-160; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n2_$name.err | tail -4
}
run auto QR_X=1
run twoshot QR_PEER_FUSED=0
run nccl QR_PEER_REDUCE=0
echo "=== N=2, 1M docs, traced (auto)"; QR_TRACE=1 QR_TRACE_ROUNDS=1 timeout 200 python scripts/longrun_sharded.py 2 201 2>&1 | grep -E "trace|trees|exchange" | tail -36
echo "=== N=2, 1M docs (auto)"; timeout 200 python scripts/longrun_sharded.py 2 300 2>&1 | grep -E "trees|exchange"
