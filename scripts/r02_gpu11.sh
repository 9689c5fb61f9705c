#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_integration_boundary.py tests/test_host_cli.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r02_pytest_integ.log 2>&1
tail -12 gpurun_out/r02_pytest_integ.log
python scripts/tree_diff_probe.py 2>&1 | tail -12
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','windows','parity','reference_mode'):
    print(k, json.dumps(d.get(k))[:900])
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','us_per_launch','kernel_ms_per_tree','kernel_share_of_step')})
PY
