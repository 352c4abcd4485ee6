#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py -k "edge or l1edge or L1Edge or raw or t19 or empty or ties or unclear or windows or property" -s -q --timeout 600 > gpurun_out/pytest_r2d.log 2>&1; grep -E "passed|failed|Error|assert|T=19|raw-HR" gpurun_out/pytest_r2d.log | head -30
python scripts/shift_loss_probe.py 65536 sobel_l1_mix 5; python scripts/shift_loss_probe.py 65536 l1 5
