cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
for per in 2 20 200; do
  timeout 900 python bench.py --workload sixteen_armadillos --steps 20 --warmup 5 --no-cpu-baseline --clock-period-ms $per > $O/r02g_${per}.json 2> $O/r02g_${per}.err
  python -c "
import json; d=json.load(open('$O/r02g_${per}.json')); f=d['e2e']['two_frames_in_flight']; print('period $per ms: value', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'in flight', round(f['value']), round(f['ms_per_step'],4), 'clock samples', d['clocks']['samples'])"
done
