#!/bin/bash
# Round 2, GPU call 31 (8 GPUs): bench at N=8 (weak cfg2 + strong cfg5 1e7); every rank under its own timeout
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02p_topo_n8.txt 2>&1
( time timeout 840 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02p_bench_n8.json 2> gpurun_out/r02p_bench_n8.err )
cat gpurun_out/r02p_bench_n8.json | cut -c1-6000; tail -n 6 gpurun_out/r02p_bench_n8.err
