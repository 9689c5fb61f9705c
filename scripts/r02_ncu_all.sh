#!/bin/bash
# round-2 ncu evidence: --set full captures of every kernel of the hot path (late, deep tree of the config-2 workload).
# The reports stay on the box (/tmp: 10-60 MB each); the per-launch summaries (scripts/ncu_summary.py) come back.
mkdir -p gpurun_out/prof
export QR_NO_PDL=1
N="--set full --import-source on --clock-control none"
ncu $N -k regex:hist_limb --launch-skip 3000 -c 34 -o /tmp/r02_hist_full -f python scripts/longrun.py 150 > /tmp/ncu_a.log 2>&1; tail -1 /tmp/ncu_a.log
python scripts/ncu_summary.py /tmp/r02_hist_full.ncu-rep gpurun_out/prof/r02_hist_full_summary "hist_limb_kernel, 34 consecutive launches (one deep tree, ~tree 100 of a config-2 run), ncu --set full --clock-control none, round 2"
ncu $N -k regex:'route_kernel' --launch-skip 3000 -c 12 -o /tmp/r02_route_full -f python scripts/longrun.py 150 > /tmp/ncu_b.log 2>&1; tail -1 /tmp/ncu_b.log
python scripts/ncu_summary.py /tmp/r02_route_full.ncu-rep gpurun_out/prof/r02_route_full_summary "route_kernel, 12 consecutive launches of a deep tree, ncu --set full, round 2"
ncu $N -k regex:'scan_pub_kernel' --launch-skip 3000 -c 12 -o /tmp/r02_scan_full -f python scripts/longrun.py 150 > /tmp/ncu_b2.log 2>&1; tail -1 /tmp/ncu_b2.log
python scripts/ncu_summary.py /tmp/r02_scan_full.ncu-rep gpurun_out/prof/r02_scan_pub_full_summary "scan_pub_kernel, 12 consecutive launches of a deep tree, ncu --set full, round 2"
for k in lambda_kernel rank_kernel leaf_node_kernel; do
  ncu $N -k regex:$k --launch-skip 100 -c 2 -o /tmp/r02_$k -f python scripts/longrun.py 150 > /tmp/ncu_c.log 2>&1; tail -1 /tmp/ncu_c.log
  python scripts/ncu_summary.py /tmp/r02_$k.ncu-rep gpurun_out/prof/r02_${k}_full_summary "$k, launches 100-101 of a config-2 run, ncu --set full, round 2"
done
ncu $N -k regex:'score_codes_kernel' -c 3 -o /tmp/r02_score_full -f python scripts/score_probe.py --n 200000 > /tmp/ncu_d.log 2>&1; tail -2 /tmp/ncu_d.log
python scripts/ncu_summary.py /tmp/r02_score_full.ncu-rep gpurun_out/prof/r02_score_codes_full_summary "score_codes_kernel, 200k docs x 700 features x 5000 trees, ncu --set full, round 2"
ncu -i /tmp/r02_score_full.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_pick.py > gpurun_out/prof/r02_score_codes_counters.txt
ncu -i /tmp/r02_lambda_kernel.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_pick.py > gpurun_out/prof/r02_lambda_counters.txt
ncu -i /tmp/r02_hist_full.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_pick.py > gpurun_out/prof/r02_hist_counters.txt
ls -la gpurun_out/prof
