#!/bin/bash
# Round 2, GPU call 38 (8 GPUs): the full BASELINE configs[4] graph (5e7 vertices / 4e8 edges) split over 8 GPUs.
# Host memory first: every rank plans its partition from the shared edge list (about 27 GB per rank at this size).
mkdir -p gpurun_out
free -g > gpurun_out/r02t_free_n8.txt; nproc >> gpurun_out/r02t_free_n8.txt; cat gpurun_out/r02t_free_n8.txt
avail=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
if [ "$avail" -lt 330 ]; then echo "only $avail GB of host memory available: the 8-rank plan of the 4e8-edge graph is not attempted"; exit 0; fi
( time timeout 1300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --strong-workload cfg5_kuramoto_er_5e7 --no-cpu-baseline > gpurun_out/r02t_bench_n8_cfg5_full.json 2> gpurun_out/r02t_bench_n8_cfg5_full.err )
cut -c1-200 gpurun_out/r02t_bench_n8_cfg5_full.json; tail -n 8 gpurun_out/r02t_bench_n8_cfg5_full.err; free -g
