cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "in_flight or pipelined or host_mirror" 2>&1 | tail -5
for w in sixteen_armadillos big_ben_clock two_armadillos; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $O/r02f_bench_${w}_n1.json 2> $O/r02f_bench_${w}_n1.err
  echo "== $w rc=$?"; tail -3 $O/r02f_bench_${w}_n1.err; python -c "
import json; d=json.load(open('$O/r02f_bench_${w}_n1.json')); print(' value', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'in flight', d['e2e'].get('two_frames_in_flight'))"
done
