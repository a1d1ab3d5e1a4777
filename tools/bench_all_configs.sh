# device-resident + e2e numbers of every BASELINE configuration (parity-test shapes; the bench line proper is c2)
for c in c1 c2 c3l0 c3l1 c3l2 c4 c5; do
  echo -n "$c: "; timeout 300 python bench.py --config $c --steps 96 --warmup 8 --no-cpu-baseline 2>gpurun_out/bench_$c.err | python -c "
import json,sys
t=sys.stdin.read().strip()
if not t: print('FAILED'); sys.exit(0)
d=json.loads(t.splitlines()[-1]); r=d['roofline']
print(round(d['value']), 'clouds/s', round(d['ms_per_step']*1e3,1), 'us/step  e2e', round(d['e2e']['value']), ' kernels/step', d['run']['kernels_per_step'], ' hbm_frac', round(r['frac'],4), r['kernel'], json.dumps({k:v['us_per_launch'] for k,v in d['roofline_detail'].items()}))
"; done
