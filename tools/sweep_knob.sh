#!/bin/bash
# Tuning helper (GPU box): rebuild the library with different values of one compile-time knob and time the configs.
#   KNOB=BVHT_SUB_CH SWEEP="0 1 2" tools/sweep_knob.sh
KNOB=${KNOB:-BVHT_SUB_CH}
for v in ${SWEEP:-0 1 2}; do
  env $KNOB=$v python -m bvhtracer_b200.build --force > /dev/null 2>&1
  echo "== $KNOB=$v"
  python tools/quick_bench.py ${CASES:-two_armadillos sixteen_armadillos sixteen_armadillos_f30 trippy_teapots big_ben_clock} 2>&1 | grep -E "${MODES:-strict-accel|fast-accel}" | cut -c1-200
done
python -m bvhtracer_b200.build --force > /dev/null 2>&1
