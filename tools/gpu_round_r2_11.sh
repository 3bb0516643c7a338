#!/bin/bash
# Round 2, GPU call 11: software-pipelined jagged kernel (index loads of the next iteration in flight during the gathers)
mkdir -p gpurun_out
B="ND_B200_KERNEL=jag,ND_B200_JAG_WINDOW=128"
MODES="auto:;u2w48_pk:$B,ND_B200_PACK_P=1;u2w32_pk:$B,ND_B200_JAG_WPS=32,ND_B200_PACK_P=1;u4w32_pk:$B,ND_B200_JAG_U=4,ND_B200_JAG_WPS=32,ND_B200_PACK_P=1;u4w24_pk:$B,ND_B200_JAG_U=4,ND_B200_JAG_WPS=24,ND_B200_PACK_P=1;u2w48:$B;u2w32:$B,ND_B200_JAG_WPS=32;u4w32:$B,ND_B200_JAG_U=4,ND_B200_JAG_WPS=32;u2w32_w32_pk:ND_B200_KERNEL=jag,ND_B200_JAG_WPS=32,ND_B200_PACK_P=1"
timeout 900 python tools/bench_configs.py cfg2 cfg2kura cfg3 --check "--modes=$MODES" > gpurun_out/r02_11_sweep.jsonl 2> gpurun_out/r02_11_sweep.err
timeout 300 python tools/bench_configs.py cfg4 cfg1 --check "--modes=auto:;jag48:ND_B200_KERNEL=jag,ND_B200_JAG_WPS=48;fused:ND_B200_KERNEL=fused" >> gpurun_out/r02_11_sweep.jsonl 2>> gpurun_out/r02_11_sweep.err
python tools/fmt_bench.py < gpurun_out/r02_11_sweep.jsonl
tail -n 5 gpurun_out/r02_11_sweep.err
