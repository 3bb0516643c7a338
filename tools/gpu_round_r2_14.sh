#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_14_bench_n1.json 2> gpurun_out/r02_14_bench_n1.err )
cat gpurun_out/r02_14_bench_n1.json; tail -n 8 gpurun_out/r02_14_bench_n1.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_14_bench_ref_n1.json 2>> gpurun_out/r02_14_bench_n1.err )
cat gpurun_out/r02_14_bench_ref_n1.json
