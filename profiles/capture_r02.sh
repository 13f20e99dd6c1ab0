#!/bin/bash
# Round-2 evidence at HEAD, one box (run under gpurun): bench line, ncu launch list of the bench command, ncu --set full of one
# launch sequence AT THE BENCH'S BATCH SIZE (1024 frames), floor probe, single-frame latency, the other configurations.
set -u
touch realtime_urdf_filter_b200/*.so
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_final_bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ruf_raster_filter|ruf_setup_bin|ruf_pose|ruf_tile_info" -s 12 -c 4 \
    -o gpurun_out/r02_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-frames 8 --ring 1 > gpurun_out/r02_final_ncu.log 2>&1
python profiles/floor_probe.py 2>&1 | grep -E "^(empty|C1|C2)" > gpurun_out/r02_floor_probe.txt
python profiles/latency.py > gpurun_out/r02_latency.txt 2>&1
for d in 0 3; do echo "RUF_DIRECT=$d RUF_CLUSTER=0 RUF_FINE_MESHLETS=0 RUF_BG_CACHE=0 (round-2 first half: copy nodes, throughput kernels)" >> gpurun_out/r02_latency.txt
  RUF_DIRECT=$d RUF_CLUSTER=0 RUF_FINE_MESHLETS=0 RUF_BG_CACHE=0 python profiles/latency.py 2>&1 | grep pinned >> gpurun_out/r02_latency.txt; done
python profiles/latency_breakdown.py > gpurun_out/r02_latency_breakdown.txt 2>&1
# kernel durations of the single-frame graph (warm caches, as a 30 Hz caller sees them)
ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --cache-control none --csv --log-file gpurun_out/r02_single_frame_launches.csv \
    python profiles/single_frame.py 12 > /dev/null 2>&1
python profiles/small_batch_probe.py > gpurun_out/r02_small_batch_probe.txt 2>&1
for c in c1 c3 c5; do python bench.py --config $c --cpu-seconds 5 > gpurun_out/r02_bench_$c.json 2>/dev/null; done
ls -la gpurun_out/r02_*
