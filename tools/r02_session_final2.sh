cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
for w in cube two_armadillos trippy_teapots; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 > $O/r02r_bench_${w}_n1.json 2> $O/r02r_bench_${w}_n1.err
  echo "== $w rc=$?"; python -c "
import json; d=json.load(open('$O/r02r_bench_${w}_n1.json')); r=d['roofline']; print(' value', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'frac', r['frac'], 'k1', r['launch_ms'], 'frame', r['frame_ms'], 'launches', d['gpu_launches'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_cover --launch-skip 3 --launch-count 1 -f -o $O/r02_k7_full python tools/prof_one.py strict-accel 6 sixteen_armadillos > $O/r02_k7_full.log 2>&1; tail -2 $O/r02_k7_full.log
