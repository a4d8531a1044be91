#!/bin/bash
# Turn the raw ncu output of tools/capture_profiles.sh (gpurun_out/r2f/) into the tracked summaries under profiles/.
set -e
D=gpurun_out/r2f
python tools/summarize_profiles.py launches $D/launches_bench.csv profiles/r2_launches_bench_summary.txt "ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 : python bench.py --steps 40 --warmup 3 --no-extras   (agh-map-16384, round 2, final kernel)" > /dev/null
python tools/summarize_profiles.py kernel $D/agh.ncu-rep profiles/r2_cat_world_kernel_agh-map-16384.txt "ncu --set full --clock-control none --import-source on -k regex:cat_world -s 160 -c 2 : python bench.py --steps 60 --warmup 3 --no-extras (launches inside the timed rotation, agh-map-16384, final round-2 kernel)" 49152 > /dev/null
python tools/summarize_profiles.py kernel $D/sq.ncu-rep profiles/r2_cat_world_kernel_squarinth-4096.txt "ncu --set full --clock-control none --import-source on -k regex:cat_world -s 400 -c 2 : python bench.py --steps 60 --warmup 3 --no-extras --workload squarinth-4096 (final round-2 kernel)" 12288 > /dev/null
python tools/summarize_profiles.py traffic profiles/traffic.json squarinth-4096=$D/sq.ncu-rep agh-map-16384=$D/agh.ncu-rep
gzip -9 -c $D/launches_bench.csv > profiles/r2_launches_bench.csv.gz
python tools/summarize_gae.py
