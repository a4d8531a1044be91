mkdir -p gpurun_out/r2
(python -m pytest tests/test_parity_gpu.py tests/test_env_gpu.py -x -q 2>&1 | tail -25) > gpurun_out/r2/t1.log
for mw in 32 16 8; do
echo "== max warps per CTA $mw"
CAT_MAX_WARPS_PER_CTA=$mw python tools/prof_step.py --map agh-map --free 1 --worlds 16384
CAT_MAX_WARPS_PER_CTA=$mw python tools/prof_step.py --map squarinth --worlds 4096
done > gpurun_out/r2/p1.log 2>&1
python tools/raster_stats.py > gpurun_out/r2/stats1.log 2>&1
cat gpurun_out/r2/t1.log gpurun_out/r2/p1.log gpurun_out/r2/stats1.log
