#!/bin/bash
# A/B of (library build, run-time setting) pairs in one GPU session. Usage: gpu_variants_env.sh "NAME VAR=x [VAR=y]" ...; NAME "base" = in-tree.
mkdir -p gpurun_out
: > gpurun_out/variants_env.log
for spec in "$@"; do
	set -- $spec
	v=$1; shift
	lib=lumen_b200/csrc/variants/$v/liblumen_b200.so
	[ "$v" = base ] && lib=lumen_b200/csrc/liblumen_b200.so
	echo -n "$v $*: " | tee -a gpurun_out/variants_env.log
	env LMB_LIB=$PWD/$lib "$@" python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-bdpt --no-config5 --no-config4 2> gpurun_out/variant_env.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('VALUE', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'stages', {k:round(v,2) for k,v in d['roofline']['stage_ms'].items()})
" | tee -a gpurun_out/variants_env.log
done
