# One GPU-box pass that regenerates what a round is judged on (run through gpurun; outputs in gpurun_out/).
#   gpurun --timeout 2400 -- 'bash tools/gpu_round_check.sh <tag>'
tag=${1:-rX}
bash tools/gpu_pass1.sh $tag                      # pytest -m gpu, smoke(), bench.py (C2), bench.py --config c3
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
bash tools/bench_all_configs.sh > gpurun_out/all_configs_$tag.txt 2>&1; cat gpurun_out/all_configs_$tag.txt   # every BASELINE configuration
bash tools/gpu_prof_r2.sh $tag                    # launch list + ncu --set full over one ring of the timed kernels (+ sources)
python tools/bench_c5_chamfer.py > gpurun_out/bench_chamfer_$tag.txt 2>&1; python tools/bench_knn.py > gpurun_out/bench_knn_$tag.txt 2>&1; python tools/bench_fps.py > gpurun_out/bench_fps_$tag.txt 2>&1
# multi-GPU (gpurun --gpus N): bash tools/gpu_n8_proof.sh N <tag> 10   and   bash tools/gpu_pass_n2.sh N <tag> 3
