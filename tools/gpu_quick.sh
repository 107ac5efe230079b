#!/bin/bash
# quick GPU iteration: parity tests + short bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_err.log | tee gpurun_out/bench_quick.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('VALUE', d['value'], 'Mrays/s  ms/step', d['ms_per_step'], 'spp/s', d['spp_per_s'], 'e2e', d['e2e']['value'])
print('stages', d['roofline']['stage_ms'], 'nodes/ray', d["roofline"]["algorithmic"]["nodes_per_ray"], 'tris/ray', d["roofline"]["algorithmic"]["tris_per_ray"], 'achieved GB/s', d['roofline']['achieved'])
print('clocks', d['clocks'], 'launches', d['gpu_launches'])
"
tail -3 gpurun_out/bench_err.log
