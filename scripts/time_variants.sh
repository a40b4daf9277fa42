#!/bin/bash
# Builds K_A variants on the GPU box and times them (scratch).
for cfg in "2 4" "3 2" "2 2" "3 4"; do
  set -- $cfg
  touch photobundle_b200/csrc/k_step.cu
  make -C photobundle_b200/csrc EXTRA="-DK_STEP_MIN_CTAS=$1 -DK_STEP_GROUP=$2" > /dev/null 2>&1
  echo "== MIN_CTAS=$1 GROUP=$2"
  python scripts/quick_gpu.py 2>&1 | grep -E "K1 cfg2: 100|solve cfg3|residual exact" | tail -3
done
touch photobundle_b200/csrc/k_step.cu
make -C photobundle_b200/csrc > /dev/null 2>&1
