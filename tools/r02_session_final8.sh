cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
for w in sixteen_armadillos cube two_armadillos trippy_teapots big_ben_clock; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $O/r02x_bench_${w}_n1.json 2> $O/r02x_bench_${w}_n1.err
  echo "== $w rc=$?"; tail -2 $O/r02x_bench_${w}_n1.err | cut -c1-300; python -c "
import json; d=json.load(open('$O/r02x_bench_${w}_n1.json')); r=d['roofline']; f=d['e2e']['two_frames_in_flight']; print(' value', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'in flight', round(f['value']), round(f['ms_per_step'],4), 'frac', round(r['frac'],3), 'launches', d['gpu_launches'])"
done
timeout 200 python tools/e2e_timeline.py sixteen_armadillos -1 1:0:1 > $O/r02x_timeline_c3.txt 2>&1; head -8 $O/r02x_timeline_c3.txt
