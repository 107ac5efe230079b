#!/bin/bash
# short bench only (no tests); prints the headline numbers
mkdir -p gpurun_out
python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_err.log | tee gpurun_out/bench_quick.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('VALUE', round(d['value'],1), 'Mrays/s  ms/step', round(d['ms_per_step'],2), 'stages', {k:round(v,2) for k,v in d['roofline']['stage_ms'].items()}, 'nodes/ray', round(d["roofline"]["algorithmic"]["nodes_per_ray"],2), 'tris/ray', round(d["roofline"]["algorithmic"]["tris_per_ray"],2))
"
tail -3 gpurun_out/bench_err.log
