#!/bin/bash
# Round 2, GPU call 41: programmatic dependent launch between the stage kernels of the captured RK4 graph
mkdir -p gpurun_out
MODES="plain:;pdl:ND_B200_RK4_PDL=1;fused_plain:ND_B200_KERNEL=fused;fused_pdl:ND_B200_KERNEL=fused,ND_B200_RK4_PDL=1"
timeout 600 python tools/bench_configs.py cfg4 cfg1 cfg2 --check "--modes=$MODES" > gpurun_out/r02w_sweep_rk4_pdl.jsonl 2> gpurun_out/r02w.err
python tools/fmt_bench.py < gpurun_out/r02w_sweep_rk4_pdl.jsonl; tail -n 5 gpurun_out/r02w.err
( time ND_B200_RK4_PDL=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zzzz_full_size.py -m gpu -x -q -k "rk4 or cfg4" ) > gpurun_out/r02w_pytest_rk4_pdl.log 2>&1
tail -n 5 gpurun_out/r02w_pytest_rk4_pdl.log
