#!/bin/bash
# Round 2, GPU call 4: first run of the streamed jagged kernel (rhs_js_kernel): parity + timing of its launch shapes
mkdir -p gpurun_out
MODES="fused:ND_B200_KERNEL=fused;js88:ND_B200_KERNEL=js;js88_pk:ND_B200_KERNEL=js,ND_B200_PACK_P=1;js48_pk:ND_B200_KERNEL=js,ND_B200_JS_U=4,ND_B200_PACK_P=1;js84_pk:ND_B200_KERNEL=js,ND_B200_JS_NST=4,ND_B200_PACK_P=1;js44_pk:ND_B200_KERNEL=js,ND_B200_JS_U=4,ND_B200_JS_NST=4,ND_B200_PACK_P=1;js84:ND_B200_KERNEL=js,ND_B200_JS_NST=4;js44:ND_B200_KERNEL=js,ND_B200_JS_U=4,ND_B200_JS_NST=4;js88_pk_w8:ND_B200_KERNEL=js,ND_B200_JS_WPS=8,ND_B200_PACK_P=1;js88_pk_w32:ND_B200_KERNEL=js,ND_B200_JS_WPS=32,ND_B200_PACK_P=1"
timeout 900 python tools/bench_configs.py cfg2 cfg2kura cfg3 cfg1 --check "--modes=$MODES" > gpurun_out/r02_4_sweep_js.jsonl 2> gpurun_out/r02_4_sweep_js.err
python tools/fmt_bench.py < gpurun_out/r02_4_sweep_js.jsonl
tail -5 gpurun_out/r02_4_sweep_js.err
