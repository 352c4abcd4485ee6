#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-scene-infer --no-side-tf32 > gpurun_out/dpab_$name.json 2> gpurun_out/dpab_$name.err
  python -c "
import json; z=json.load(open('gpurun_out/dpab_$name.json')); print('$name', 'ms/step', round(z['ms_per_step'],3), 'patches/s', round(z['value'],1), 'e2e', round(z['e2e']['value'],1))"
}
run overlap PV_X=1
run nooverlap PV_DP_OVERLAP=0
run overlap_ch2 NCCL_MAX_NCHANNELS=2
run overlap_ch1_ll NCCL_MAX_NCHANNELS=1 NCCL_PROTO=LL
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-scene-infer --no-side-tf32 > gpurun_out/dpab_single.json 2>/dev/null
python -c "
import json; z=json.load(open('gpurun_out/dpab_single.json')); print('single', 'ms/step', round(z['ms_per_step'],3), 'patches/s', round(z['value'],1))"
