#!/bin/bash
# Round 2, GPU call 30 (2 GPUs): 2-GPU parity tests, bench at N=2 (weak cfg2 + strong cfg5 1e7), reference arm at N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02o_topo_n2.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q -k "two_gpu or multi" ) > gpurun_out/r02o_pytest_2gpu.log 2>&1
tail -n 8 gpurun_out/r02o_pytest_2gpu.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02o_bench_n2.json 2> gpurun_out/r02o_bench_n2.err )
cat gpurun_out/r02o_bench_n2.json; tail -n 6 gpurun_out/r02o_bench_n2.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02o_bench_ref_n2.json 2>> gpurun_out/r02o_bench_n2.err )
cat gpurun_out/r02o_bench_ref_n2.json
