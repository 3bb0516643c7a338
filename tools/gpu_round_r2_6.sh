#!/bin/bash
# Round 2, GPU call 6: deep-unroll variants of the jagged kernel (U = 4 / 8 columns per iteration), window 128
mkdir -p gpurun_out
B="ND_B200_KERNEL=jag,ND_B200_JAG_WINDOW=128"
MODES="u2w48_pk:$B,ND_B200_PACK_P=1;u4w32_pk:$B,ND_B200_JAG_U=4,ND_B200_JAG_WPS=32,ND_B200_PACK_P=1;u4w48_pk:$B,ND_B200_JAG_U=4,ND_B200_JAG_WPS=48,ND_B200_PACK_P=1;u8w24_pk:$B,ND_B200_JAG_U=8,ND_B200_JAG_WPS=24,ND_B200_PACK_P=1;u8w32_pk:$B,ND_B200_JAG_U=8,ND_B200_JAG_WPS=32,ND_B200_PACK_P=1;u8w48_pk:$B,ND_B200_JAG_U=8,ND_B200_JAG_WPS=48,ND_B200_PACK_P=1;u4w32:$B,ND_B200_JAG_U=4,ND_B200_JAG_WPS=32;u8w24:$B,ND_B200_JAG_U=8,ND_B200_JAG_WPS=24;u8w32:$B,ND_B200_JAG_U=8,ND_B200_JAG_WPS=32"
timeout 900 python tools/bench_configs.py cfg2 cfg2kura cfg3 --check "--modes=$MODES" > gpurun_out/r02_6_sweep_jagdeep.jsonl 2> gpurun_out/r02_6_sweep_jagdeep.err
python tools/fmt_bench.py < gpurun_out/r02_6_sweep_jagdeep.jsonl
tail -n 5 gpurun_out/r02_6_sweep_jagdeep.err
