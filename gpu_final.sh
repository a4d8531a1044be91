mkdir -p gpurun_out/r2
python bench.py > gpurun_out/r2/bench_1gpu.json 2> gpurun_out/r2/bench_1gpu.err
python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/r2/bench_ref_1gpu.json 2> gpurun_out/r2/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2/launches_bench.csv python bench.py --steps 40 --warmup 3 --no-extras > gpurun_out/r2/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cat_world -s 160 -c 2 -f -o gpurun_out/r2/prof_bench_agh python bench.py --steps 60 --warmup 3 --no-extras > gpurun_out/r2/ncu_bench_agh.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cat_world -s 400 -c 2 -f -o gpurun_out/r2/prof_bench_sq python bench.py --steps 60 --warmup 3 --no-extras --workload squarinth-4096 > gpurun_out/r2/ncu_bench_sq.log 2>&1
tail -c 300 gpurun_out/r2/bench_1gpu.err; ls -la gpurun_out/r2/
