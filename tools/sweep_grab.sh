#!/bin/bash
# GPU box: rebuild with BVHT_GRAB = 1, 2, 4, 8 and time the frames that are most / least sensitive to the work counter
for g in 1 2 4 8; do
  BVHT_GRAB=$g python -m bvhtracer_b200.build --force > /dev/null 2>&1
  echo "== BVHT_GRAB=$g"
  PYTHONPATH=. python tools/floor_probe.py 2>&1 | grep "strict-accel" | cut -c1-50
  python tools/quick_bench.py two_armadillos sixteen_armadillos sixteen_armadillos_f30 trippy_teapots big_ben_clock 2>&1 | grep "strict-accel" | cut -c1-90
done
python -m bvhtracer_b200.build --force > /dev/null 2>&1
