#!/bin/bash
# A/B of run-time settings of the in-tree library in one GPU session. Usage: gpu_env_ab.sh "VAR=a" "VAR=b VAR2=c" ...  ("-" = defaults)
mkdir -p gpurun_out
: > gpurun_out/env_ab.log
for v in "$@"; do
	[ "$v" = "-" ] && v="LMB_NOOP=1"
	echo -n "$v: " | tee -a gpurun_out/env_ab.log
	env $v python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-bdpt --no-config5 --no-config4 2> gpurun_out/env_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('VALUE', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'stages', {k:round(v,2) for k,v in d['roofline']['stage_ms'].items()})
" | tee -a gpurun_out/env_ab.log
	tail -2 gpurun_out/env_ab.err
done
