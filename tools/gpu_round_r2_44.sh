#!/bin/bash
# Round 2, GPU call 44 (4 GPUs): bench.py as the driver launches it at N=4 (defaults: weak cfg2 + the full configs[4] graph)
mkdir -p gpurun_out
free -g | head -2
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02z_bench_n4_defaults.json 2> gpurun_out/r02z_bench_n4_defaults.err )
cut -c1-250 gpurun_out/r02z_bench_n4_defaults.json; tail -n 6 gpurun_out/r02z_bench_n4_defaults.err
