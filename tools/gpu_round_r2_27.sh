#!/bin/bash
# Round 2, GPU call 27: L2 prefetch of the entry streams (cp.async.bulk.prefetch.L2, one per slice, fixed distance ahead)
mkdir -p gpurun_out
J="ND_B200_KERNEL=jagb,ND_B200_JAG_WINDOW=128"
MODES="auto:;pf300k:ND_B200_PF_DIST=300000;pf600k:ND_B200_PF_DIST=600000;pf1200k:ND_B200_PF_DIST=1200000;pf2400k:ND_B200_PF_DIST=2400000;pf1200k_w48:ND_B200_PF_DIST=1200000,ND_B200_JAG_WPS=48;pf1200k_u4w32:ND_B200_PF_DIST=1200000,ND_B200_JAG_U=4;b4w48_pf:$J,ND_B200_JAGA_CH=4,ND_B200_JAGA_WPS=48,ND_B200_PF_DIST=1200000;b6w40_pf:$J,ND_B200_JAGA_CH=6,ND_B200_JAGA_WPS=40,ND_B200_PF_DIST=1200000;b8w32_pf:$J,ND_B200_JAGA_WPS=32,ND_B200_PF_DIST=1200000;b8w32_pf2400k:$J,ND_B200_JAGA_WPS=32,ND_B200_PF_DIST=2400000"
timeout 900 python tools/bench_configs.py cfg2 cfg2kura cfg2nop --check "--modes=$MODES" > gpurun_out/r02l_sweep_l2_prefetch.jsonl 2> gpurun_out/r02l.err
python tools/fmt_bench.py < gpurun_out/r02l_sweep_l2_prefetch.jsonl
tail -n 5 gpurun_out/r02l.err
