# first GPU pass of a change: tests, smoke, bench (N=1)
tag=${1:-r2a}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','reps','ms_per_step_min','ms_per_step_p90','parity_check','allreduce_check')})
    print('e2e', d['e2e']['value'], 'single', d['single_launch'] and d['single_launch']['value'])
    print("roofline", {k:d["roofline"][k] for k in ("kernel","bound","frac","hbm_frac","fp32_frac")}); print("dropin", d.get("dropin"))
    print({k:v['us_per_launch'] for k,v in d['roofline_detail'].items()})
except Exception as e: print('no bench line', e)
PY
timeout 600 python bench.py --config c3 --steps 40 --warmup 8 > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err; echo "bench c3 rc=$?"; tail -2 gpurun_out/bench_c3_$tag.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c3_$tag.json').read().strip().splitlines()[-1])
    print('c3', {k:d[k] for k in ('value','ms_per_step','reps','parity_check')}, 'e2e', d['e2e']['value'], d['roofline']['bound'], d['roofline']['frac'], d['cpu_baseline']['value'])
except Exception as e: print('no c3 line', e)
PY
