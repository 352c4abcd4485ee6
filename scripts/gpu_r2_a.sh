#!/bin/bash
# round-2 visit A: full-batch golden gradients of every engine, baseline bench, ncu --set full of the loss and reduction kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_r2a.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_rows.py -k "golden or tensor_core_kernels" -s -q --timeout 900 > gpurun_out/pytest_r2a.log 2>&1
grep -E "B=128|passed|failed|Error|error" gpurun_out/pytest_r2a.log | head -20
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"
python - <<PY
import json
z=json.load(open("gpurun_out/bench_r2a.json"))
print("patches/s", round(z["value"],1), "ms/step", round(z["ms_per_step"],3), "e2e", round(z["e2e"]["value"],1), z["clocks"])
print(z["roofline"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shift_loss -c 3 -f -o gpurun_out/prof_r2_shiftloss_b128 python scripts/shift_loss_probe.py 128 l1 1 > gpurun_out/ncu_sl128.log 2>&1; tail -2 gpurun_out/ncu_sl128.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shift_loss -c 3 -f -o gpurun_out/prof_r2_shiftloss_b65536 python scripts/shift_loss_probe.py 65536 l1 1 > gpurun_out/ncu_sl64k.log 2>&1; tail -2 gpurun_out/ncu_sl64k.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shift_loss -c 3 -f -o gpurun_out/prof_r2_shiftloss_edge_b65536 python scripts/shift_loss_probe.py 65536 sobel_l1_mix 1 > gpurun_out/ncu_sle64k.log 2>&1; tail -2 gpurun_out/ncu_sle64k.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deferred_reduce -s 2 -c 2 -f -o gpurun_out/prof_r2_deferred_reduce python scripts/profile_fwd.py tf32 3 > gpurun_out/ncu_dr.log 2>&1; tail -2 gpurun_out/ncu_dr.log
python scripts/shift_loss_probe.py 65536 l1 5; python scripts/shift_loss_probe.py 65536 sobel_l1_mix 5; python scripts/shift_loss_probe.py 128 l1 20
