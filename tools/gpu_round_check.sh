set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r1c.log
python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_r1c.json 2> gpurun_out/bench_ref_r1c.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_r1c.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"fps_reg|knn_group|chamfer_small|hard_mask" -s 4 -c 4 -o gpurun_out/prof_c2_r1c python tools/prof_kernels.py --config c2 --reps 1 > gpurun_out/ncu_c2_r1c.log 2>&1; echo "ncu full rc=$?"
python -c "import __graft_entry__ as g; g.smoke()"
