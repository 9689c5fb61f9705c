#!/bin/bash
# single-GPU evidence pass: GPU tests, smoke, the bench line, the launch list of the bench command, ncu of the REFERENCE-mode kernels
mkdir -p gpurun_out/prof
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err; cut -c1-400 gpurun_out/r02_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --settle 30 --late-window 0 --e2e-trees 2 --no-cpu-baseline --reference-mode-trees 0 > gpurun_out/r02_bench_ncu.log 2>&1
tail -2 gpurun_out/r02_bench_ncu.log | cut -c1-200
python scripts/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/prof/r02_launches_summary.txt 2>&1; head -30 gpurun_out/prof/r02_launches_summary.txt
N="--set full --import-source on --clock-control none"
for k in hist_exact_walk_kernel hist_exact_list_kernel ordered_squares_resolve_kernel leaf_exact_kernel; do
  timeout 600 ncu $N -k regex:$k --launch-skip 6 -c 4 -o /tmp/r02_$k -f python scripts/refmode_probe.py 3 > /tmp/ncu_c.log 2>&1; tail -1 /tmp/ncu_c.log
  python scripts/ncu_summary.py /tmp/r02_$k.ncu-rep gpurun_out/prof/r02_${k}_full_summary "$k, 4 launches of the second tree of a config-2 run in QR_HIST_REFERENCE mode, ncu --set full, round 2" | tail -2
done
timeout 300 python scripts/refmode_probe.py 40 2>&1 | tail -4
