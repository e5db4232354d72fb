cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "in_flight" 2>&1 | tail -1
for w in big_ben_clock sixteen_armadillos trippy_teapots; do
  timeout 100 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t); f=d['e2e']['two_frames_in_flight']; print('$w value', round(d['value']), 'e2e', round(d['e2e']['value']), 'in flight', round(f['value']), round(f['ms_per_step'],4), f['last_frame_equals_render'])
except Exception: print('$w:', t[-600:])"
done
