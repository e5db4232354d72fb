cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02t_pytest.txt 2>&1; tail -3 $O/r02t_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for w in sixteen_armadillos big_ben_clock; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 > $O/r02t_bench_${w}_n1.json 2> $O/r02t_bench_${w}_n1.err
  echo "== $w rc=$?"; python -c "
import json; d=json.load(open('$O/r02t_bench_${w}_n1.json')); r=d['roofline']; print(' value', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'frac', r['frac'], 'k1', r['launch_ms'], 'frame', r['frame_ms'], 'alg', r.get('algorithmic_speedup'), 'launches', d['gpu_launches'])"
done
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02t_bench_reference.json 2>&1; cut -c1-160 $O/r02t_bench_reference.json
timeout 300 python tools/e2e_timeline.py sixteen_armadillos -1 > $O/r02t_timeline_c3.txt 2>&1; grep -E "device frame|kernels done|copy done" $O/r02t_timeline_c3.txt
(for tool in memcheck racecheck; do echo "#### compute-sanitizer --tool $tool python tools/sanitize_run.py"; timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -v "^$" | tail -12; done) > $O/r02t_sanitizer.txt 2>&1; cat $O/r02t_sanitizer.txt | tail -30
