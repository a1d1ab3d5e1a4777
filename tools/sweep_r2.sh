python tools/quick_step.py --path dataflow --check
python tools/quick_step.py --path single --check
python tools/quick_step.py --path dataflow --parts c
python tools/bench_c5_chamfer.py
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
