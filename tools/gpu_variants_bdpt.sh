#!/bin/bash
# A/B of library builds on the BDPT table (tools/bdpt_table.py). Usage: gpu_variants_bdpt.sh NAME [NAME ...]; "base" = the in-tree library.
mkdir -p gpurun_out
: > gpurun_out/variants_bdpt.log
for v in "$@"; do
	lib=lumen_b200/csrc/variants/$v/liblumen_b200.so
	[ "$v" = base ] && lib=lumen_b200/csrc/liblumen_b200.so
	echo -n "$v: " | tee -a gpurun_out/variants_bdpt.log
	LMB_LIB=$PWD/$lib python tools/bdpt_table.py 2> gpurun_out/variant_bdpt_$v.err | python -c "
import json,sys
print(' '.join(f\"{r['scene']} {r['ms_per_frame']:.2f}\" for r in map(json.loads, sys.stdin)))
" | tee -a gpurun_out/variants_bdpt.log
done
