#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the 2-GPU parity tests and bench.py under torchrun at N ranks.
# usage: gpu_multi.sh N [steps]
N=${1:-2}; STEPS=${2:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv,noheader
python -m pytest tests/test_gpu_comm.py tests/test_headless_cli.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 3 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N', d['n_gpus'], 'VALUE', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'film_check', d['film_check'])
print('config4', {k: d['config4'][k] for k in ('value','ms_per_step','spp_per_s')}, d['config4']['e2e'])
"
tail -5 gpurun_out/bench_n$N.err
