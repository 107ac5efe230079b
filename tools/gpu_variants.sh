#!/bin/bash
# A/B of library builds (tools/build_variant.sh) in one GPU session. Usage: gpu_variants.sh NAME [NAME ...]; "base" = the in-tree library.
mkdir -p gpurun_out
: > gpurun_out/variants.log
for v in "$@"; do
	lib=lumen_b200/csrc/variants/$v/liblumen_b200.so
	[ "$v" = base ] && lib=lumen_b200/csrc/liblumen_b200.so
	echo -n "$v: " | tee -a gpurun_out/variants.log
	LMB_LIB=$PWD/$lib python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-bdpt --no-config5 --no-config4 2> gpurun_out/variant_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
a=d['roofline']['algorithmic']
print('VALUE', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'stages', {k:round(v,2) for k,v in d['roofline']['stage_ms'].items()}, 'nodes/ray', round(a['nodes_per_ray'],2), 'tris/ray', round(a['tris_per_ray'],2))
" | tee -a gpurun_out/variants.log
	grep "k_trace profile" gpurun_out/variant_$v.err | tail -1 | tee -a gpurun_out/variants.log
done
