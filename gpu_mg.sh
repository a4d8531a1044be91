mkdir -p gpurun_out/r2
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/r2/bench_${N}gpu.json 2> gpurun_out/r2/bench_${N}gpu.err
tail -c 400 gpurun_out/r2/bench_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 64 --workload mappo-agh-map > gpurun_out/r2/train_${N}gpu.json 2> gpurun_out/r2/train_${N}gpu.err
tail -c 600 gpurun_out/r2/train_${N}gpu.err; head -c 1500 gpurun_out/r2/train_${N}gpu.json
