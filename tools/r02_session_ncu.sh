cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_primary --launch-skip 3 --launch-count 1 -f -o $O/r02_k1_full python tools/prof_one.py strict-accel 6 sixteen_armadillos > $O/r02_k1_full.log 2>&1; tail -2 $O/r02_k1_full.log
timeout 300 python tools/stats_dump.py sixteen_armadillos 2 2>/dev/null | tail -1 > $O/r02_stats_c3f2.json; cut -c1-300 $O/r02_stats_c3f2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02_bench_under_ncu.log 2>&1; tail -c 300 $O/r02_bench_under_ncu.log; wc -l $O/r02_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_primary --launch-skip 3 --launch-count 1 -f -o $O/r02_k1_full_c5 python tools/prof_one.py strict-accel 6 big_ben_clock > $O/r02_k1_full_c5.log 2>&1; tail -2 $O/r02_k1_full_c5.log
ls -la $O/*.ncu-rep
