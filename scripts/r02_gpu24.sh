#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
python bench.py --e2e-trees 400 --no-cpu-baseline --reference-mode-trees 0 > gpurun_out/r02_bench_try.json 2> gpurun_out/r02_bench_try.err; tail -2 gpurun_out/r02_bench_try.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_try.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "hist us/launch", d["roofline"]["us_per_launch"], "kernel ms/tree", d["roofline"]["kernel_ms_per_tree"], d["roofline"]["phase_ms_per_tree_with_sync"], d["windows"])
PY
