#!/bin/bash
# compute-sanitizer over the hot path: memcheck + racecheck on one GPU, memcheck on a
# 2-GPU sharded run when two devices are visible.  Logs go to gpurun_out/ and are copied to profiles/.
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_target.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok" gpurun_out/r02_sanitizer_$tool.log
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 900 compute-sanitizer --tool memcheck --target-processes all --print-limit 20 python scripts/longrun_sharded.py 2 3 30000 > gpurun_out/r02_sanitizer_memcheck_2gpu.log 2>&1
  echo "== memcheck 2 GPUs: rc=$?"; grep -E "ERROR SUMMARY" gpurun_out/r02_sanitizer_memcheck_2gpu.log | sort | uniq -c
fi
