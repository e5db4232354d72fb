#!/bin/bash
# GPU session of round 2: parity tests, work counters + ncu counters per workload (roofline calibration / traffic), option A/B,
# bench lines of every BASELINE config.  Everything lands in gpurun_out/r02c_*.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02c_pytest.txt 2>&1; tail -3 $O/r02c_pytest.txt
CASES="sixteen_armadillos:2 sixteen_armadillos:15 two_armadillos:1 trippy_teapots:10 big_ben_clock:3 cube:1"
rm -f $O/r02c_stats.jsonl
for c in $CASES; do
  w=${c%%:*}; f=${c##*:}
  python tools/stats_dump.py $w $f 2>>$O/r02c_stats.err | tail -1 >> $O/r02c_stats.jsonl
  ncu --metrics smsp__thread_inst_executed.sum,smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum \
      --clock-control none -k regex:"trace_primary|classify_fill|raster" -c 60 --csv --log-file $O/r02c_ncu_${w}_${f}.csv python tools/stats_dump.py $w $f > /dev/null 2>>$O/r02c_stats.err
done
cat $O/r02c_stats.jsonl | cut -c1-600
python tools/cover_ab.py > $O/r02c_cover_ab.txt 2>&1; cat $O/r02c_cover_ab.txt
for w in sixteen_armadillos cube two_armadillos trippy_teapots big_ben_clock; do
  python bench.py --workload $w --steps 20 --warmup 5 > $O/r02c_bench_${w}.json 2> $O/r02c_bench_${w}.err
  echo "== $w rc=$?"; cut -c1-300 $O/r02c_bench_${w}.json; tail -2 $O/r02c_bench_${w}.err
done
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02c_bench_reference.json 2>&1; cut -c1-300 $O/r02c_bench_reference.json
