# refresh of the ring capture only (the timed kernels over one whole ring), summarised on the box
tag=${1:-r2}
timeout 900 ncu --set full --clock-control none -k regex:"cloud_step|chamfer_warp32" -s 48 -c 48 -o /tmp/prof_ring_$tag -f python tools/prof_ring.py --config c2 --rings 2 > gpurun_out/ncu_ring_$tag.log 2>&1; echo "ncu ring rc=$?"
python tools/ncu_ring_json.py /tmp/prof_ring_$tag.ncu-rep --kernel cloud_step > gpurun_out/ncu_ring_group_$tag.json
python tools/ncu_ring_json.py /tmp/prof_ring_$tag.ncu-rep --kernel chamfer_warp32 > gpurun_out/ncu_ring_chamfer_$tag.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cloud_step|chamfer_warp32" -s 48 -c 2 -o gpurun_out/prof_ring_src_$tag -f python tools/prof_ring.py --config c2 --rings 2 > gpurun_out/ncu_ring_src_$tag.log 2>&1; echo "ncu src rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 20 --warmup 5 --reps 2 --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1; echo "ncu launches rc=$?"
