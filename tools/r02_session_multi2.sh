# usage: r02_session_multi2.sh <tag> <N> [<N> ...]   (run on a box with at least max(N) GPUs)
cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
O=gpurun_out
T=$1; shift
for w in sixteen_armadillos big_ben_clock; do
  for n in "$@"; do
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --workload $w --gpus $n --steps 20 --warmup 5 > $O/${T}_bench_${w}_n$n.json 2> $O/${T}_bench_${w}_n$n.err
    echo "== $w N=$n rc=$?"; tail -1 $O/${T}_bench_${w}_n$n.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(' value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'launches', d['gpu_launches'], 'ok', d['sharded_frame_equals_single_gpu'], d['gathered_device_frame_equals_single_gpu'], 'roof', d['roofline'].get('frac'))
except Exception as e: print('parse error', e)"
    grep -v "OMP_NUM_THREADS\|^\*\*\*" $O/${T}_bench_${w}_n$n.err | tail -2
  done
done
