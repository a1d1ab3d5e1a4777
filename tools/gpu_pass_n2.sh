N=${1:-2}; tag=${2:-r2}; runs=${3:-2}
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_dist_${tag}_n$N.log 2>&1; echo "pytest dist rc=$?"; tail -2 gpurun_out/pytest_dist_${tag}_n$N.log
for i in $(seq 1 $runs); do
GM3D_BENCH_WATCHDOG_S=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+i)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${tag}_n${N}_run$i.json 2> gpurun_out/bench_${tag}_n${N}_run$i.err
echo "rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_${tag}_n${N}_run$i.json').read().strip().splitlines()[-1]); print('N=$N run $i K=20:', round(d['value']), round(d['ms_per_step']*1e3,2), 'us/step', [round(x*1e3,2) for x in d['rank_median_ms_per_step']], d['allreduce_check'], d['parity_check'], 'e2e', round(d['e2e']['value']))"
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${tag}_n1_k20.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_${tag}_n1_k20.json').read().strip().splitlines()[-1]); print('N=1 K=20:', round(d['value']), round(d['ms_per_step']*1e3,2), 'us/step reps', d['reps'])"
