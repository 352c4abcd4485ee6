#!/bin/bash
# round-2 multi-GPU visit (run with gpurun --gpus N): data-parallel check, weak and strong scaling of the default bench, config 5
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR scripts/dp_check.py > gpurun_out/dp_check_${N}gpu.log 2>&1; echo "dp_check rc=$?"; tail -3 gpurun_out/dp_check_${N}gpu.log
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_weak_${N}gpu.json 2> gpurun_out/bench_r02_weak_${N}gpu.err; echo "weak rc=$?"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-scene-infer --scaling strong > gpurun_out/bench_r02_strong_${N}gpu.json 2> gpurun_out/bench_r02_strong_${N}gpu.err; echo "strong rc=$?"
timeout 900 $TR bench.py --gpus $N --cfg cfg/p16t9c85r24f64.cfg --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline --no-scene-infer --no-side-tf32 > gpurun_out/bench_r02_cfg5_fp32_${N}gpu.json 2> gpurun_out/bench_r02_cfg5_${N}gpu.err; echo "cfg5 rc=$?"
python - <<PY
import json
for f in ("weak", "strong", "cfg5_fp32"):
    try:
        z = json.load(open(f"gpurun_out/bench_r02_{f}_${N}gpu.json"))
        print(f, z["n_gpus"], "patches/s", round(z["value"], 1), "ms/step", round(z["ms_per_step"], 3), "e2e", round(z["e2e"]["value"], 1), z["config"]["global_batch"], z["dtype"], z["scaling"],
              z.get("single_pass_tf32") and round(z["single_pass_tf32"]["value"], 1), z.get("scene_infer") and round(z["scene_infer"]["value"], 1))
    except Exception as e:
        print(f, "failed", e)
PY
