N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02d_bench_${N}gpu.json 2> gpurun_out/r02d_bench_${N}gpu.err
tail -c 300 gpurun_out/r02d_bench_${N}gpu.err
