#!/bin/bash
# weak-scaling point at N GPUs (run with gpurun --gpus N): bash scripts/gpu_r2_scale.sh N
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_weak_${N}gpu.json 2> gpurun_out/bench_r02_weak_${N}gpu.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_weak_${N}gpu.json 2> gpurun_out/bench_r02_weak_${N}gpu.err
fi
python -c "
import json; z=json.load(open('gpurun_out/bench_r02_weak_${N}gpu.json'))
print('N', z['n_gpus'], 'patches/s', round(z['value'],1), 'ms/step', round(z['ms_per_step'],3), 'e2e', round(z['e2e']['value'],1), 'tf32', round(z['single_pass_tf32']['value'],1), 'scenes/s', round(z['scene_infer']['value'],1))"
