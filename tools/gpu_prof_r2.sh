tag=${1:-r2}
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 20 --warmup 5 --reps 2 --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1; echo "ncu launches rc=$?"
# --set full over one whole ring (24 group + 24 Chamfer launches) of the timed kernels: DRAM bytes complete over the ring.
# The report is summarised on the box (it is larger than what travels back).
timeout 900 ncu --set full --clock-control none -k regex:"cloud_step|chamfer_warp32" -s 48 -c 48 -o /tmp/prof_ring_$tag -f python tools/prof_ring.py --config c2 --rings 2 > gpurun_out/ncu_ring_$tag.log 2>&1; echo "ncu ring rc=$?"
python tools/ncu_ring_json.py /tmp/prof_ring_$tag.ncu-rep --kernel cloud_step > gpurun_out/ncu_ring_group_$tag.json
python tools/ncu_ring_json.py /tmp/prof_ring_$tag.ncu-rep --kernel chamfer_warp32 > gpurun_out/ncu_ring_chamfer_$tag.json
# one launch of each with source (per-line tables)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cloud_step|chamfer_warp32" -s 48 -c 2 -o gpurun_out/prof_ring_src_$tag -f python tools/prof_ring.py --config c2 --rings 2 > gpurun_out/ncu_ring_src_$tag.log 2>&1; echo "ncu src rc=$?"
# stand-alone Chamfer at the scaling shape (C5 shard: 39 296 patch pairs)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chamfer_warp32 -s 1 -c 1 -o gpurun_out/prof_chamfer_c5_$tag -f python tools/prof_kernels.py --config c5 --reps 2 --only chamfer_fused > gpurun_out/ncu_chamfer_c5_$tag.log 2>&1; echo "ncu chamfer c5 rc=$?"
du -sh gpurun_out
