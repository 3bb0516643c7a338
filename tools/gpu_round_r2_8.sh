#!/bin/bash
# Round 2, GPU call 8: tile kernel with the skewed value staging (one pad slot per row), live and packed parameters
mkdir -p gpurun_out
MODES="fused:ND_B200_KERNEL=fused;fused_pk:ND_B200_KERNEL=fused,ND_B200_PACK_P=1"
timeout 900 python tools/bench_configs.py cfg2 cfg2nop cfg2kura cfg3 cfg1 cfg4 --check "--modes=$MODES" > gpurun_out/r02_8_sweep_skew.jsonl 2> gpurun_out/r02_8_sweep_skew.err
python tools/fmt_bench.py < gpurun_out/r02_8_sweep_skew.jsonl
tail -n 5 gpurun_out/r02_8_sweep_skew.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rhs_fused -s 8 -c 1 -f -o gpurun_out/r02_8_fused_pk_cfg2 python tools/bench_configs.py cfg2 --quick "--modes=fp:ND_B200_KERNEL=fused,ND_B200_PACK_P=1" > gpurun_out/r02_8_ncu.log 2>&1
tail -n 2 gpurun_out/r02_8_ncu.log
