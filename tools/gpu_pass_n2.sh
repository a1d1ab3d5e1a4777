N=${1:-2}; tag=${2:-r2}
timeout 300 python -m pytest tests/test_gpu_dist.py tests/test_gpu_parity.py -m gpu -x -q -k "dist or peer or two_gpus" > gpurun_out/pytest_dist_${tag}_n$N.log 2>&1; echo "pytest dist rc=$?"; tail -3 gpurun_out/pytest_dist_${tag}_n$N.log
for mode in auto peer_step nccl none; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --collective $mode > gpurun_out/bench_${tag}_n${N}_$mode.json 2> gpurun_out/bench_${tag}_n${N}_$mode.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_${tag}_n${N}_$mode.json').read().strip().splitlines()[-1]); print('N=$N $mode K=20:', round(d['value']), round(d['ms_per_step']*1e3,2), 'us/step', [round(x*1e3,2) for x in d['rank_median_ms_per_step']], d['allreduce_check'], d['parity_check'])"
done
