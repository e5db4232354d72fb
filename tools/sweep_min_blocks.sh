#!/bin/bash
# Tuning helper (GPU box): rebuild the library with different register budgets and time the headline config.
for mb in ${SWEEP:-4 5 6 8}; do
  BVHT_MIN_BLOCKS=$mb python -m bvhtracer_b200.build --force > /dev/null 2>&1
  echo "== BVHT_MIN_BLOCKS=$mb"
  python tools/quick_bench.py ${CASES:-two_armadillos sixteen_armadillos sixteen_armadillos_f30 big_ben_clock} 2>&1 | grep -E "strict-accel|strict-brute" | cut -c1-110
done
python -m bvhtracer_b200.build --force > /dev/null 2>&1
