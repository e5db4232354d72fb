cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
(for tool in memcheck racecheck; do echo "#### compute-sanitizer --tool $tool python tools/sanitize_run.py"; timeout 600 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -v "^$" | tail -13; done) > $O/r02w_sanitizer.txt 2>&1; grep -c "ok" $O/r02w_sanitizer.txt; grep "SUMMARY" $O/r02w_sanitizer.txt
for w in trippy_teapots trippy_teapots trippy_teapots; do
  timeout 100 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t); f=d['e2e']['two_frames_in_flight']; print('$w in flight', round(f['value']), f['last_frame_equals_render'])
except Exception: print('$w:', t[-600:])"
done
for w in sixteen_armadillos cube two_armadillos trippy_teapots big_ben_clock; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 5 > $O/r02w_bench_${w}_n1.json 2> $O/r02w_bench_${w}_n1.err
  echo "== $w rc=$?"; python -c "
import json; d=json.load(open('$O/r02w_bench_${w}_n1.json')); r=d['roofline']; f=d['e2e']['two_frames_in_flight']; print(' value', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'in flight', round(f['value']), round(f['ms_per_step'],4), 'frac', round(r['frac'],3), 'k1', round(r['launch_ms'],4), 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline']['value'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02w_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02w_bench_under_ncu.log 2>&1
