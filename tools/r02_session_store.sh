cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
L=bvhtracer_b200/lib
cp $L/libbvht_cuda.so /tmp/default.so
for v in st0 st1 st2 st0; do
  cp $L/variants/libbvht_cuda_$v.so $L/libbvht_cuda.so
  for w in sixteen_armadillos; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$v $w value', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'k1 warm', round(r['launch_ms'],4))"
  done
done
cp /tmp/default.so $L/libbvht_cuda.so
