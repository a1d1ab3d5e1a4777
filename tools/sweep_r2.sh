for nl in 6 8 12 16 24; do GM3D_RING_LANES=$nl python tools/quick_step.py --path dataflow; done
GM3D_RING_LANES=8 GM3D_CD_WARPS=8 python tools/quick_step.py --path dataflow
python tools/quick_step.py --path single
for c in c1 c4 c3l2; do python tools/quick_step.py --path dataflow --config $c --check; python tools/quick_step.py --path single --config $c; done
