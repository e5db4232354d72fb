cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
for w in sixteen_armadillos cube two_armadillos trippy_teapots big_ben_clock; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 > $O/r02p_bench_${w}_n1.json 2> $O/r02p_bench_${w}_n1.err
  echo "== $w rc=$?"; python -c "
import json; d=json.load(open('$O/r02p_bench_${w}_n1.json')); r=d['roofline']; print(' value', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'frac', r['frac'], 'alg_speedup', r.get('algorithmic_speedup'), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['single_thread']['value'], 'launches', d['gpu_launches'])"
  tail -2 $O/r02p_bench_${w}_n1.err
done
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02p_bench_reference.json 2>&1; cut -c1-200 $O/r02p_bench_reference.json
timeout 300 python tools/e2e_timeline.py sixteen_armadillos -1 1:0:1 > $O/r02p_timeline_c3.txt 2>&1
