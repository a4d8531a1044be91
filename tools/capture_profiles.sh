#!/bin/bash
# Run on the GPU box (gpurun -- bash tools/capture_profiles.sh): the bench line, the ncu launch list and the ncu --set full
# captures that tools/refresh_profiles.sh turns into the summaries under profiles/.  ncu runs only after the plain run exited 0.
mkdir -p gpurun_out/r2f
python bench.py --steps 60 --warmup 3 --no-extras > gpurun_out/r2f/bench_noextras.json 2> gpurun_out/r2f/bench_noextras.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2f/launches_bench.csv python bench.py --steps 40 --warmup 3 --no-extras > gpurun_out/r2f/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cat_world -s 160 -c 2 -f -o gpurun_out/r2f/agh python bench.py --steps 60 --warmup 3 --no-extras > gpurun_out/r2f/ncu_agh.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cat_world -s 400 -c 2 -f -o gpurun_out/r2f/sq python bench.py --steps 60 --warmup 3 --no-extras --workload squarinth-4096 > gpurun_out/r2f/ncu_sq.log 2>&1
for T in 256 1024; do
ncu --set full --clock-control none -k regex:"cat_gae|cat_adv" -s 6 -c 2 -f -o gpurun_out/r2f/gae_T$T python tools/prof_gae.py --T $T --iters 4 > gpurun_out/r2f/ncu_gae_$T.log 2>&1
done
ls -la gpurun_out/r2f/
