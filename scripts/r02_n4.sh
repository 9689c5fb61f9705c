#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
QR_TEST_WORLD=$N timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x -k "auto" > gpurun_out/r02_pytest_mgpu$N.log 2>&1
tail -3 gpurun_out/r02_pytest_mgpu$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_n${N}_b.json 2> gpurun_out/r02_bench_c2_n${N}_b.err
tail -2 gpurun_out/r02_bench_c2_n${N}_b.err | cut -c1-300; grep '^{' gpurun_out/r02_bench_c2_n${N}_b.json | cut -c1-300
QR_KTRACE=1 timeout 300 python scripts/longrun_sharded.py $N 150 2>&1 | grep -v "^\[ktrace\] host" | cut -c1-330 | tail -8
