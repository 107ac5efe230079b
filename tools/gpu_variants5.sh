#!/bin/bash
# A/B of library builds on the headline AND on config 5 (10 M triangles, 2^24 incoherent rays). Usage: gpu_variants5.sh NAME [NAME ...]
mkdir -p gpurun_out
: > gpurun_out/variants5.log
for v in "$@"; do
	lib=lumen_b200/csrc/variants/$v/liblumen_b200.so
	[ "$v" = base ] && lib=lumen_b200/csrc/liblumen_b200.so
	echo -n "$v: " | tee -a gpurun_out/variants5.log
	LMB_LIB=$PWD/$lib python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-bdpt --no-config4 2> gpurun_out/variant_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
c=d['roofline_config5']
print('VALUE', round(d['value'],1), 'stages', {k:round(v,2) for k,v in d['roofline']['stage_ms'].items()}, 'config5', round(c['mrays_per_s'],1), 'sorted', round(c['sorted']['mrays_per_s'],1))
" | tee -a gpurun_out/variants5.log
done
