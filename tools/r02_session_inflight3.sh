cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in sixteen_armadillos big_ben_clock two_armadillos trippy_teapots cube; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $O/r02f_bench_${w}_n1.json 2> $O/r02f_bench_${w}_n1.err
  echo "== $w rc=$?"; tail -3 $O/r02f_bench_${w}_n1.err; python -c "
import json; d=json.load(open('$O/r02f_bench_${w}_n1.json')); f=d['e2e']['two_frames_in_flight']; print(' value', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'in flight', round(f['value']), round(f['ms_per_step'],4), 'flush', round(f['l2_flush_ms_per_step'],4), f['last_frame_equals_render'], 'launches', d['gpu_launches'])"
done
