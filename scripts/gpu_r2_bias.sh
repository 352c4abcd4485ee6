#!/bin/bash
# bring-up of the bias MMA of the x3 fused forward: selftest with the default tile strides and with LBO / SBO swapped, then the bench
for d in "128,256" "256,128"; do
  echo "== PV_X3_BIAS_DESC=$d"; PV_X3_BIAS_DESC=$d timeout 200 python scripts/selftest.py 2>&1 | grep -E "x3 fused|rc "
done
