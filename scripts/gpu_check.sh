#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench (both arms), and an ncu launch list of one train step.
# Usage (from the repo root, on the GPU box): bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_${TAG}.txt 2>&1
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_${TAG}.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.log
tail -40 gpurun_out/pytest_${TAG}.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
cat gpurun_out/bench_${TAG}.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
cat gpurun_out/bench_ref_${TAG}.json
# launch list of one steady-state train step (3 warm-up steps skipped); numbers under ncu are never bench values
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-330} -c ${NCU_COUNT:-115} --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${TAG}.log 2>&1
echo "ncu rc=$?"
