#!/bin/bash
# Round 2, GPU call 20: (1) gather paths other than the LSU (TEX, LDGSTS, TMA bulk copies), (2) rhs_jaga_kernel sweep
mkdir -p gpurun_out
timeout 300 tools/_build/path_gather_bench > gpurun_out/r02f_path_gather_bench_raw.txt 2>&1
cat gpurun_out/r02f_path_gather_bench_raw.txt
MODES="auto:;jaga8w48:ND_B200_KERNEL=jaga,ND_B200_JAG_WINDOW=128;jaga8w32:ND_B200_KERNEL=jaga,ND_B200_JAG_WINDOW=128,ND_B200_JAGA_WPS=32;jaga4w64:ND_B200_KERNEL=jaga,ND_B200_JAG_WINDOW=128,ND_B200_JAGA_CH=4,ND_B200_JAGA_WPS=64;jaga4w48:ND_B200_KERNEL=jaga,ND_B200_JAG_WINDOW=128,ND_B200_JAGA_CH=4;jaga16w24:ND_B200_KERNEL=jaga,ND_B200_JAG_WINDOW=128,ND_B200_JAGA_CH=16;jaga8w48_win32:ND_B200_KERNEL=jaga,ND_B200_JAG_WINDOW=32"
timeout 900 python tools/bench_configs.py cfg2 cfg2kura cfg2nop cfg3 --check "--modes=$MODES" > gpurun_out/r02f_sweep_jaga.jsonl 2> gpurun_out/r02f_sweep_jaga.err
python tools/fmt_bench.py < gpurun_out/r02f_sweep_jaga.jsonl
tail -n 5 gpurun_out/r02f_sweep_jaga.err
MODES5="auto:;jaga8w48:ND_B200_KERNEL=jaga,ND_B200_JAG_WINDOW=128;jaga4w64:ND_B200_KERNEL=jaga,ND_B200_JAG_WINDOW=128,ND_B200_JAGA_CH=4,ND_B200_JAGA_WPS=64"
timeout 600 python tools/bench_configs.py cfg5s --check "--modes=$MODES5" > gpurun_out/r02f_sweep_jaga_cfg5s.jsonl 2>> gpurun_out/r02f_sweep_jaga.err
python tools/fmt_bench.py < gpurun_out/r02f_sweep_jaga_cfg5s.jsonl
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rhs_jaga -s 8 -c 1 -f -o gpurun_out/r02f_jaga8w48_cfg2 python tools/bench_configs.py cfg2 --quick "--modes=j:ND_B200_KERNEL=jaga,ND_B200_JAG_WINDOW=128" > gpurun_out/r02f_ncu.log 2>&1
tail -n 2 gpurun_out/r02f_ncu.log
