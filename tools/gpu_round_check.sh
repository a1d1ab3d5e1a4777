# One GPU-box pass that produces everything a round is judged on (run through gpurun; outputs in gpurun_out/).
#   bash tools/gpu_round_check.sh <tag>
tag=${1:-rX}
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
timeout 600 python bench.py --no-overlap --no-cpu-baseline > gpurun_out/bench_nooverlap_$tag.json 2> gpurun_out/bench_nooverlap_$tag.err; echo "bench no-overlap rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
# every BASELINE configuration (parity-test shapes; c2 is the bench line)
bash tools/bench_all_configs.sh > gpurun_out/all_configs_$tag.txt 2>&1; cat gpurun_out/all_configs_$tag.txt
# launch list of the same command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 48 --warmup 24 --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1; echo "ncu launches rc=$?"
# one full-set capture of the dominant kernel, and of the two-phase kNN kernel at the scaling shape
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cloud_step" -s 1 -c 1 -o gpurun_out/prof_cloud_step_$tag python tools/prof_kernels.py --config c2 --reps 2 --only cloud_step > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"knn_large" -s 1 -c 1 -o gpurun_out/prof_knn_large_$tag python tools/prof_kernels.py --config c5 --reps 2 --only knn_group > gpurun_out/ncu_knn_large_$tag.log 2>&1; echo "ncu knn_large rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fps_reg" -s 1 -c 1 -o gpurun_out/prof_fps_c5_$tag python tools/prof_kernels.py --config c5 --reps 2 --only fps > gpurun_out/ncu_fps_c5_$tag.log 2>&1; echo "ncu fps rc=$?"
./tools/ubench/fp32_pipes > gpurun_out/fp32_pipes_$tag.txt 2>&1
python tools/bench_knn.py > gpurun_out/bench_knn_$tag.txt 2>&1; python tools/bench_fps.py > gpurun_out/bench_fps_$tag.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
