#!/bin/bash
mkdir -p gpurun_out
MODES="fused:ND_B200_KERNEL=fused;fused_pk:ND_B200_KERNEL=fused,ND_B200_PACK_P=1;fused_nocr:ND_B200_KERNEL=fused,ND_B200_NO_COMPACT=1"
timeout 900 python tools/bench_configs.py cfg2 cfg2kura cfg3 --check "--modes=$MODES" > gpurun_out/r02_10_sweep.jsonl 2> gpurun_out/r02_10_sweep.err
python tools/fmt_bench.py < gpurun_out/r02_10_sweep.jsonl
tail -n 5 gpurun_out/r02_10_sweep.err
