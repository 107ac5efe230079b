#!/bin/bash
# ncu only: launch list + one full capture of kernels matching $1 (default k_trace). Outputs in gpurun_out/.
PAT=${1:-k_trace}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 4 2 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$PAT" -s ${2:-1} -c ${3:-2} -f -o gpurun_out/prof python tools/ncu_target.py 4 1 | tail -1
