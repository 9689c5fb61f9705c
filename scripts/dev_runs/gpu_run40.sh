#!/bin/bash
# 2-GPU: sharded-training parity (peer-memory exchange and NCCL) + bench at N=2 for the three exchange paths
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m | head -6
timeout 500 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_mgpu.log 2>&1
tail -25 gpurun_out/pytest_mgpu.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 50 --e2e-trees 200 > gpurun_out/bench_n2_$name.json 2> gpurun_out/bench_n2_$name.err
  echo "== $name"; tail -1 gpurun_out/bench_n2_$name.json | cut -c1-160; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n2_$name.err | tail -4
}
run peer QR_PEER_REDUCE=1
run nccl QR_PEER_REDUCE=0
run old QR_PEER_REDUCE=0 QR_COMM_3PASS=1
