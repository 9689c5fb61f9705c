#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
for n in 1 2; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_n$n.log 2> gpurun_out/bench_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 > gpurun_out/bench_n$n.log 2> gpurun_out/bench_n$n.err
  fi
  tail -1 gpurun_out/bench_n$n.log | cut -c1-200; tail -3 gpurun_out/bench_n$n.err
done
