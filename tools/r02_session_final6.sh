cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02v_bench_reference.json 2> $O/r02v_bench_reference.err; tail -c 400 $O/r02v_bench_reference.json
for w in sixteen_armadillos cube two_armadillos trippy_teapots big_ben_clock; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 5 > $O/r02v_bench_${w}_n1.json 2> $O/r02v_bench_${w}_n1.err
  echo "== $w rc=$?"; python -c "
import json; d=json.load(open('$O/r02v_bench_${w}_n1.json')); r=d['roofline']; f=d['e2e']['two_frames_in_flight']; print(' value', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'in flight', round(f['value']), 'frac', round(r['frac'],3), 'k1', round(r['launch_ms'],4), 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline']['value'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02v_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02v_bench_under_ncu.log 2>&1
timeout 200 python tools/e2e_timeline.py sixteen_armadillos -1 1:0:1 > $O/r02v_timeline_c3.txt 2>&1
