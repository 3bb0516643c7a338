#!/bin/bash
# Round 2, GPU call 39: bench.py with its defaults (what the driver runs) at N=1
mkdir -p gpurun_out
free -g | head -2
( time timeout 1500 python bench.py > gpurun_out/r02u_bench_n1_defaults.json 2> gpurun_out/r02u_bench_n1_defaults.err )
cut -c1-250 gpurun_out/r02u_bench_n1_defaults.json; tail -n 6 gpurun_out/r02u_bench_n1_defaults.err
