#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py tests/test_host_sharding.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r02_pytest_mgpu.log 2>&1
tail -8 gpurun_out/r02_pytest_mgpu.log
timeout 300 python scripts/longrun_sharded.py 2 300 2>&1 | tail -4
timeout 600 compute-sanitizer --tool memcheck --target-processes all --print-limit 20 python scripts/longrun_sharded.py 2 3 30000 > gpurun_out/r02_sanitizer_memcheck_2gpu.log 2>&1
echo "== memcheck 2 GPUs: rc=$?"; grep -E "ERROR SUMMARY" gpurun_out/r02_sanitizer_memcheck_2gpu.log | sort | uniq -c
