#!/bin/bash
# Builds liblumen_b200.so with extra nvcc flags into lumen_b200/csrc/variants/NAME/ (git-ignored *.so, ships with gpurun) for
# A/B runs: LMB_LIB=lumen_b200/csrc/variants/NAME/liblumen_b200.so python bench.py ...   Usage: build_variant.sh NAME "-DFOO=1 ..."
set -e
name=$1; shift
src=$(cd "$(dirname "$0")/../lumen_b200/csrc" && pwd)
out=$src/variants/$name
mkdir -p "$out"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC -I$src/../../include --expt-relaxed-constexpr"
pids=()
for f in capi lbvh ploc wide_bvh wavefront bdpt post testhooks comm; do
	$NVCC $FLAGS "$@" -c "$src/$f.cu" -o "$out/$f.o" & pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$out/liblumen_b200.so" "$out"/*.o -lcudart -ldl
rm -f "$out"/*.o
echo "built $out/liblumen_b200.so ($*)"
