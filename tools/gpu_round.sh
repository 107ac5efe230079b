#!/bin/bash
# One GPU session: parity tests, bench (ours + reference arm), ncu launch list and one full capture of k_trace (all 8 launches
# of one batch). Outputs in gpurun_out/; summarise here with tools/ncu_summary.py and copy into profiles/. gpurun brings back at most
# 64 MiB: the two .ncu-rep files of 16-frame batches are ~30 MB each -- run the last line as a second call if the merge is refused.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json
tail -5 gpurun_out/bench_err.log
python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 16 2 | tail -2
timeout 900 ncu --set full --clock-control none -k regex:"^k_trace$" -s 0 -c 8 -f -o gpurun_out/prof_trace python tools/ncu_target.py 16 1 | tail -2
timeout 900 ncu --set full --clock-control none -k regex:"k_shade|k_classify|k_connect|k_miss" -s 0 -c 10 -f -o gpurun_out/prof_shade python tools/ncu_target.py 16 1 | tail -2
# BDPT row: per-scene table, launch list and one full capture of the pair kernels
python tools/bdpt_table.py --json gpurun_out/bdpt_table.json | cut -c1-160
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bdpt_launches.csv python tools/bdpt_table.py --quick | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_bdpt_pair -s 2 -c 2 -f -o gpurun_out/prof_bdpt_pair python tools/bdpt_table.py --quick | tail -1
ls -la gpurun_out
