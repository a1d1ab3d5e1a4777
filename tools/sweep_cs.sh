# usage: bash tools/sweep_cs.sh "<warps list>" "<modes list>" -- fused-kernel step time (us) for tuning knobs
for w in ${1:-24}; do for m in ${2:-0}; do
  echo -n "warps=$w mode=$m: "; GM3D_CS_WARPS=$w GM3D_CS_MODE=$m python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2),'us')"
done; done
