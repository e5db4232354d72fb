cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02m_pytest.txt 2>&1; tail -3 $O/r02m_pytest.txt
timeout 300 python tools/e2e_timeline.py big_ben_clock -1 2>&1 | grep -v "copy done" > $O/r02m_timeline_c5.txt; cat $O/r02m_timeline_c5.txt
for w in big_ben_clock two_armadillos; do
timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $O/r02m_bench_$w.json 2> $O/r02m_bench_$w.err; python -c "
import json; d=json.load(open('$O/r02m_bench_$w.json')); print('$w value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], 'launches', d['gpu_launches'])"
done
