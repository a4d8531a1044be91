mkdir -p gpurun_out/r2
ncu --set full --clock-control none --import-source on -k regex:cat_world -s 70 -c 1 -f -o gpurun_out/r2/prof_r2_agh python tools/prof_step.py --map agh-map --free 1 --worlds 16384 --steps 60 > gpurun_out/r2/ncu_agh.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cat_world -s 70 -c 1 -f -o gpurun_out/r2/prof_r2_sq python tools/prof_step.py --map squarinth --worlds 4096 --steps 60 > gpurun_out/r2/ncu_sq.log 2>&1
tail -n 2 gpurun_out/r2/ncu_agh.log gpurun_out/r2/ncu_sq.log
