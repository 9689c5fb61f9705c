#!/bin/bash
# usage: r02_scale.sh N  — bench lines at N GPUs (configs 2, 4 and, at N=8, 3) + the sharded parity tests at world N
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_n$N.json 2> gpurun_out/r02_bench_c2_n$N.err
tail -2 gpurun_out/r02_bench_c2_n$N.err | cut -c1-300; cut -c1-400 gpurun_out/r02_bench_c2_n$N.json
$TR bench.py --config 4 --gpus $N --steps 20 --warmup 5 --e2e-trees 100 > gpurun_out/r02_bench_c4_n$N.json 2> gpurun_out/r02_bench_c4_n$N.err
tail -2 gpurun_out/r02_bench_c4_n$N.err | cut -c1-300; cut -c1-400 gpurun_out/r02_bench_c4_n$N.json
if [ "$N" = "8" ]; then
  $TR bench.py --config 3 --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_c3_n$N.json 2> gpurun_out/r02_bench_c3_n$N.err
  tail -2 gpurun_out/r02_bench_c3_n$N.err | cut -c1-300; cut -c1-400 gpurun_out/r02_bench_c3_n$N.json
  QR_TEST_WORLD=8 timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x -k "auto" > gpurun_out/r02_pytest_mgpu8.log 2>&1
  tail -4 gpurun_out/r02_pytest_mgpu8.log
fi
