#!/bin/bash
# Tuning helper (GPU box): rebuild the library with different register budgets and time the headline config.
for mb in 4 5 6 8; do
  BVHT_MIN_BLOCKS=$mb python -m bvhtracer_b200.build --force > /dev/null 2>&1
  echo "== BVHT_MIN_BLOCKS=$mb"
  python tools/quick_bench.py two_armadillos sixteen_armadillos big_ben_clock 2>&1 | grep -E "strict-accel|strict-brute"
done
python -m bvhtracer_b200.build --force > /dev/null 2>&1
