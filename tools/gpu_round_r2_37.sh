#!/bin/bash
# Round 2, GPU call 37: the full BASELINE configs[4] graph (Kuramoto on Erdos-Renyi 5e7 vertices / 4e8 edges) on ONE GPU
mkdir -p gpurun_out
free -g > gpurun_out/r02s_free.txt; nproc >> gpurun_out/r02s_free.txt
( time timeout 1500 python bench.py --steps 10 --warmup 3 --strong-workload cfg5_kuramoto_er_5e7 --no-cpu-baseline > gpurun_out/r02s_bench_n1_cfg5_full.json 2> gpurun_out/r02s_bench_n1_cfg5_full.err )
cat gpurun_out/r02s_free.txt; cut -c1-300 gpurun_out/r02s_bench_n1_cfg5_full.json; tail -n 8 gpurun_out/r02s_bench_n1_cfg5_full.err
