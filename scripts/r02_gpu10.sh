#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_gpu.log 2>&1
tail -6 gpurun_out/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','windows','parity','reference_mode','clocks','cpu_baseline'):
    print(k, json.dumps(d.get(k))[:600])
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','us_per_launch','kernel_ms_per_tree','kernel_share_of_step')})
print('scoring', d['scoring']['value'], d['scoring']['e2e'])
PY
bash scripts/r02_sanitize.sh
