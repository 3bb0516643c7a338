#!/bin/bash
# Round 2, GPU call 24: batched jagged kernel (rhs_jagb_kernel) sweep
mkdir -p gpurun_out
J="ND_B200_KERNEL=jagb,ND_B200_JAG_WINDOW=128"
MODES="auto:;b8w32:$J,ND_B200_JAGA_WPS=32;b8w40:$J,ND_B200_JAGA_WPS=40;b8w48:$J,ND_B200_JAGA_WPS=48;b6w40:$J,ND_B200_JAGA_CH=6,ND_B200_JAGA_WPS=40;b6w48:$J,ND_B200_JAGA_CH=6,ND_B200_JAGA_WPS=48;b4w48:$J,ND_B200_JAGA_CH=4,ND_B200_JAGA_WPS=48;b4w64:$J,ND_B200_JAGA_CH=4,ND_B200_JAGA_WPS=64;b16w24:$J,ND_B200_JAGA_CH=16;b8w32_win32:ND_B200_KERNEL=jagb,ND_B200_JAG_WINDOW=32,ND_B200_JAGA_WPS=32"
timeout 900 python tools/bench_configs.py cfg2 cfg2kura cfg2nop cfg3 --check "--modes=$MODES" > gpurun_out/r02i_sweep_jagb.jsonl 2> gpurun_out/r02i.err
python tools/fmt_bench.py < gpurun_out/r02i_sweep_jagb.jsonl
tail -n 5 gpurun_out/r02i.err
( time timeout 900 python -m pytest tests/test_zzzz_round2_kernels.py -m gpu -x -q ) > gpurun_out/r02i_pytest_round2_kernels.log 2>&1
tail -n 6 gpurun_out/r02i_pytest_round2_kernels.log
