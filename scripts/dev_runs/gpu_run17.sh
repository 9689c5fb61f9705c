#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'encode_kernel|score_codes' -c 12 --csv --log-file gpurun_out/score_launches.csv python scripts/score_probe.py --n 1000000 --f 700 --trees 5000 > gpurun_out/score_ncu.log 2>&1
grep -v "^==" gpurun_out/score_launches.csv | cut -d, -f5,8,9,14,15 | head -14
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_codes -s 1 -c 1 -o gpurun_out/prof_score python scripts/score_probe.py --n 200000 --f 700 --trees 5000 > gpurun_out/score_ncu2.log 2>&1
tail -2 gpurun_out/score_ncu2.log
