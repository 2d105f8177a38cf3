timeout 900 python -m pytest tests/test_gpu_dp.py -q -m gpu 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02e_bench_2gpu.json 2> gpurun_out/r02e_bench_2gpu.err
tail -c 200 gpurun_out/r02e_bench_2gpu.err
