#!/bin/bash
# Round 2, GPU call 29: the whole -m gpu suite at HEAD, then bench.py (both arms) at N=1, launch list of the bench under ncu
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02n_pytest_gpu.log 2>&1
tail -n 8 gpurun_out/r02n_pytest_gpu.log
( time timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r02n_bench_n1.json 2> gpurun_out/r02n_bench_n1.err )
cat gpurun_out/r02n_bench_n1.json; tail -n 5 gpurun_out/r02n_bench_n1.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02n_bench_ref_n1.json 2>> gpurun_out/r02n_bench_n1.err )
cat gpurun_out/r02n_bench_ref_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02n_launches_bench.csv python bench.py --steps 5 --warmup 3 > gpurun_out/r02n_bench_under_ncu.log 2>&1
tail -n 3 gpurun_out/r02n_bench_under_ncu.log
