cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
(for tool in memcheck racecheck; do echo "#### compute-sanitizer --tool $tool python tools/sanitize_run.py"; timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -v "^$" | tail -13; done) > $O/r02u_sanitizer.txt 2>&1; tail -30 $O/r02u_sanitizer.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_primary --launch-skip 3 --launch-count 1 -f -o $O/r02u_k1_full python tools/prof_one.py strict-accel 6 sixteen_armadillos > $O/r02u_k1_full.log 2>&1
timeout 300 python tools/stats_dump.py sixteen_armadillos 2 2>/dev/null | tail -1 > $O/r02u_stats_c3f2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_primary --launch-skip 3 --launch-count 1 -f -o $O/r02u_k1_full_c5 python tools/prof_one.py strict-accel 6 big_ben_clock > $O/r02u_k1_full_c5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02u_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02u_bench_under_ncu.log 2>&1
CASES="sixteen_armadillos:2 sixteen_armadillos:15 two_armadillos:1 trippy_teapots:10 big_ben_clock:3 cube:1"
rm -f $O/r02u_stats.jsonl
for c in $CASES; do
  w=${c%%:*}; f=${c##*:}
  python tools/stats_dump.py $w $f 2>/dev/null | tail -1 >> $O/r02u_stats.jsonl
  ncu --metrics smsp__thread_inst_executed.sum,smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum \
      --clock-control none -k regex:"trace_primary|classify_fill|raster|small_ops" -c 80 --csv --log-file $O/r02u_ncu_${w}_${f}.csv python tools/stats_dump.py $w $f > /dev/null 2>&1
done
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02u_bench_reference.json 2>&1
for w in sixteen_armadillos cube two_armadillos trippy_teapots big_ben_clock; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 > $O/r02u_bench_${w}_n1.json 2> $O/r02u_bench_${w}_n1.err
  echo "== $w rc=$?"; python -c "
import json; d=json.load(open('$O/r02u_bench_${w}_n1.json')); r=d['roofline']; f=d['e2e']['two_frames_in_flight']; print(' value', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'in flight', round(f['value']), 'frac', round(r['frac'],3), 'k1', round(r['launch_ms'],4), 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline']['value'])"
done
timeout 300 python tools/e2e_timeline.py sixteen_armadillos -1 1:0:1 > $O/r02u_timeline_c3.txt 2>&1
timeout 300 python tools/inflight_probe.py sixteen_armadillos 30 > $O/r02u_inflight_probe.txt 2>&1
