#!/bin/bash
# A/B sweep of k_trace's scheduling thresholds (env overrides; results do not depend on them). Usage: gpu_sweep.sh [notest]
mkdir -p gpurun_out
[ "$1" = notest ] || python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
run() {
	echo -n "REFILL=$1 TRI=$2: "
	LMB_REFILL_LANES=$1 LMB_TRI_ROUND_LANES=$2 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>> gpurun_out/bench_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('VALUE', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'stages', {k:round(v,2) for k,v in d['roofline']['stage_ms'].items()}, 'nodes/ray', round(d['roofline']['nodes_per_ray'],2))
"
}
for r in ${REFILLS:-20 24 26 28 30}; do run $r 8; done 2>&1 | tee gpurun_out/sweep.log
for t in ${TRIS:-6 10}; do run 26 $t; done 2>&1 | tee -a gpurun_out/sweep.log
tail -3 gpurun_out/bench_err.log
