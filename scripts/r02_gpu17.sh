#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
QR_KTRACE=20 timeout 300 python scripts/longrun_sharded.py $N 200 > gpurun_out/r02_ktrace17_n$N.log 2>&1; grep -v "^\[ktrace\] tree 99\|tree 49" gpurun_out/r02_ktrace17_n$N.log | tail -14
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_n$N.json 2> gpurun_out/r02_bench_c2_n$N.err
tail -2 gpurun_out/r02_bench_c2_n$N.err | cut -c1-300; cut -c1-400 gpurun_out/r02_bench_c2_n$N.json
