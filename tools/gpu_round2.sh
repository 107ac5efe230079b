#!/bin/bash
# Round-2 GPU session: parity tests, bench (ours + reference arm), ncu launch list, the calibration pass of the issue roofline
# (per-launch smsp__inst_executed.sum joined with the kernel's own counters) and one full capture of k_trace. Outputs in gpurun_out/;
# summarise HERE with tools/ncu_summary.py / tools/calibrate_ktrace.py and copy into profiles/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
if [ "$1" != "notests" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log; fi
python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json | cut -c1-600
tail -5 gpurun_out/bench_err.log
python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/bench_ref.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 16 2 | tail -2
LMB_STATS_PER_LAUNCH=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'^k_trace$' --csv --log-file gpurun_out/ktrace_inst.csv python tools/ncu_target.py 16 1 2> gpurun_out/ktrace_counters.log | tail -1
LMB_STATS_PER_LAUNCH=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'^k_trace$' --csv --log-file gpurun_out/ktrace_inst_4.csv python tools/ncu_target.py 4 1 2> gpurun_out/ktrace_counters_4.log | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_trace$" -s 0 -c 8 -f -o gpurun_out/prof_trace python tools/ncu_target.py 16 1 | tail -2
ls -la gpurun_out
