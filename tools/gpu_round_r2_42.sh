#!/bin/bash
# Round 2, GPU call 42: programmatic dependent launch as the default of the captured RK4 graph (index loads ahead of pdl_wait)
mkdir -p gpurun_out
MODES="auto:;nopdl:ND_B200_RK4_PDL=0"
timeout 900 python tools/bench_configs.py cfg4 cfg1 cfg2 cfg2kura cfg3 cfg5s --check "--modes=$MODES" > gpurun_out/r02x_sweep_rk4_pdl_default.jsonl 2> gpurun_out/r02x.err
python tools/fmt_bench.py < gpurun_out/r02x_sweep_rk4_pdl_default.jsonl; tail -n 5 gpurun_out/r02x.err
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02x_pytest_gpu.log 2>&1
tail -n 5 gpurun_out/r02x_pytest_gpu.log
