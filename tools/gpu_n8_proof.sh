# N-GPU robustness proof: R consecutive full bench runs (the driver's command line), each bounded; one log line per run
N=${1:-8}; tag=${2:-r2}; runs=${3:-10}
mkdir -p gpurun_out
log=gpurun_out/bench_n${N}_consecutive_$tag.log
: > $log
for i in $(seq 1 $runs); do
  t0=$(date +%s)
  GM3D_BENCH_WATCHDOG_S=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+i)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_${tag}_run$i.json 2> gpurun_out/bench_n${N}_${tag}_run$i.err
  rc=$?
  python - >> $log <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_${tag}_run$i.json').read().strip().splitlines()[-1])
    print('run $i rc=$rc wall=%ds' % ($(date +%s)-$t0), 'value', round(d['value']), 'us/step', round(d['ms_per_step']*1e3,2), 'reps', d['reps'], 'parity', d['parity_check'], 'allreduce', d['allreduce_check'], 'rank medians', [round(x*1e3,2) for x in d['rank_median_ms_per_step']], 'e2e', round(d['e2e']['value']), 'collective:', d['run']['collective'][:40])
except Exception as e:
    print('run $i rc=$rc NO JSON LINE', e)
PY
  tail -1 $log
done
# keep one full JSON line + stage trace, drop the rest of the per-run files
cp gpurun_out/bench_n${N}_${tag}_run1.json gpurun_out/bench_n${N}_$tag.json
grep -v Warning gpurun_out/bench_n${N}_${tag}_run1.err | grep "bench r0" > gpurun_out/bench_n${N}_${tag}_stages.txt
