cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
N=${1:-2}
T=${2:-r02k}
for w in sixteen_armadillos big_ben_clock; do
  for n in 1 2 4 8; do
    [ $n -gt $N ] && continue
    if [ $n -eq 1 ]; then
      timeout 900 python bench.py --workload $w --gpus 1 --steps 20 --warmup 5 > $O/${T}_bench_${w}_n$n.json 2> $O/${T}_bench_${w}_n$n.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --workload $w --gpus $n --steps 20 --warmup 5 > $O/${T}_bench_${w}_n$n.json 2> $O/${T}_bench_${w}_n$n.err
    fi
    echo "== $w N=$n rc=$?"; tail -1 $O/${T}_bench_${w}_n$n.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(' value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'launches', d['gpu_launches'], 'ok', d['sharded_frame_equals_single_gpu'], d['gathered_device_frame_equals_single_gpu'], 'roof', d['roofline'].get('frac'))
except Exception as e: print('parse error', e)"
    grep -v "OMP_NUM_THREADS\|^\*\*\*" $O/${T}_bench_${w}_n$n.err | tail -2
  done
done
