mkdir -p gpurun_out/r2
python -m pytest tests/test_env_gpu.py -q -x -k gae 2>&1 | tail -2
for T in 256 1024; do
ncu --set full --clock-control none -k regex:"cat_gae|cat_adv" -s 6 -c 2 -f -o gpurun_out/r2/prof_gae_T$T python tools/prof_gae.py --T $T --iters 4 > gpurun_out/r2/ncu_gae_$T.log 2>&1
done
ls -la gpurun_out/r2/*.ncu-rep
