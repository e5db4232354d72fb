cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coverage or sixteen or equivalence or pipelined" 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"raster" -c 12 --csv --log-file $O/r02_k7_times.csv python tools/prof_one.py strict-accel 6 sixteen_armadillos > /dev/null 2>&1; grep raster $O/r02_k7_times.csv | awk -F'","' '{print $5, $NF}' | tail -6
timeout 300 python tools/cover_ab.py sixteen_armadillos 2>&1 | tail -1
