cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
for kick in 0 0 0 1 1 1 1 1 1; do
  if [ $kick = 1 ]; then export BVHT_KICK=1; else unset BVHT_KICK; fi
  timeout 45 python bench.py --workload trippy_teapots --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-700 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t); f=d['e2e']['two_frames_in_flight']; print('kick $kick: ok, in flight', round(f['value']), round(f['ms_per_step'],4))
except Exception: print('kick $kick:', t)"
done
