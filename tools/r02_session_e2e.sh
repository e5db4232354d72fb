cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02i_pytest.txt 2>&1; tail -4 $O/r02i_pytest.txt
timeout 600 python tools/e2e_timeline.py sixteen_armadillos -1 12:1:2 12:1:1 16:1:2 > $O/r02i_timeline_c3.txt 2>&1; cat $O/r02i_timeline_c3.txt
timeout 300 python tools/e2e_timeline.py big_ben_clock -1 > $O/r02i_timeline_c5.txt 2>&1; cat $O/r02i_timeline_c5.txt
timeout 300 python tools/e2e_timeline.py two_armadillos -1 > $O/r02i_timeline_c2.txt 2>&1; cat $O/r02i_timeline_c2.txt
timeout 300 python tools/e2e_timeline.py cube -1 > $O/r02i_timeline_c1.txt 2>&1; cat $O/r02i_timeline_c1.txt
for w in sixteen_armadillos two_armadillos big_ben_clock; do
timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $O/r02i_bench_$w.json 2> $O/r02i_bench_$w.err; python -c "
import json; d=json.load(open('$O/r02i_bench_$w.json')); print('$w value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], d['roofline'].get('algorithmic_speedup'), 'launches', d['gpu_launches'])"
done
