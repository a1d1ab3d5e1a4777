# multi-GPU pass: dist tests + repeated bench runs at N GPUs (gpurun --gpus N)
N=${1:-2}; tag=${2:-r2}; runs=${3:-3}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_dist_${tag}_n$N.log 2>&1; echo "pytest dist rc=$?"; tail -3 gpurun_out/pytest_dist_${tag}_n$N.log
for i in $(seq 1 $runs); do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+i)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${tag}_n${N}_run$i.json 2> gpurun_out/bench_${tag}_n${N}_run$i.err
  echo "run $i rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${tag}_n${N}_run$i.json').read().strip().splitlines()[-1])
    print(round(d['value']), round(d['ms_per_step']*1e3,2), 'us/step reps', d['reps'], d['parity_check'], d['allreduce_check'], d['run']['collective'][:20], 'e2e', round(d['e2e']['value']))
except Exception as e: print('no line', e)
PY
)"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --collective nccl > gpurun_out/bench_${tag}_n${N}_nccl.json 2> gpurun_out/bench_${tag}_n${N}_nccl.err
echo "nccl rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${tag}_n${N}_nccl.json').read().strip().splitlines()[-1])
    print(round(d['value']), round(d['ms_per_step']*1e3,2), 'us/step', d['parity_check'], d['allreduce_check'], d['run']['collective'][:20])
except Exception as e: print('no line', e)
PY
)"
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${tag}_n1_k20.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_${tag}_n1_k20.json').read().strip().splitlines()[-1]); print('N=1 K=20:', round(d['value']), round(d['ms_per_step']*1e3,2), 'us/step reps', d['reps'])"
