#!/bin/bash
# Round 2, GPU call 21: persistent warps / 64-thread blocks of the jagged kernel; persistent cooperative RK4; full-size parity
mkdir -p gpurun_out
J="ND_B200_KERNEL=jag,ND_B200_JAG_WINDOW=128"
MODES="auto:;persist_u2w32:$J,ND_B200_JAG_PERSIST=1,ND_B200_JAG_WPS=32;persist_u2w48:$J,ND_B200_JAG_PERSIST=1,ND_B200_JAG_WPS=48;persist_u2w64:$J,ND_B200_JAG_PERSIST=1,ND_B200_JAG_WPS=64;persist_u4w32:$J,ND_B200_JAG_PERSIST=1,ND_B200_JAG_U=4,ND_B200_JAG_WPS=32;persist_u4w48:$J,ND_B200_JAG_PERSIST=1,ND_B200_JAG_U=4,ND_B200_JAG_WPS=48;b64w32:$J,ND_B200_JAG_BLOCK=64,ND_B200_JAG_WPS=32;b64w48:$J,ND_B200_JAG_BLOCK=64,ND_B200_JAG_WPS=48;b64w64:$J,ND_B200_JAG_BLOCK=64,ND_B200_JAG_WPS=64"
timeout 900 python tools/bench_configs.py cfg2 cfg2kura cfg2nop --check "--modes=$MODES" > gpurun_out/r02g_sweep_persist.jsonl 2> gpurun_out/r02g_sweep_persist.err
python tools/fmt_bench.py < gpurun_out/r02g_sweep_persist.jsonl
tail -n 5 gpurun_out/r02g_sweep_persist.err
MODESR="auto:;nocoop:ND_B200_RK4_COOP=0;jag_coop:ND_B200_KERNEL=jag;jag_nocoop:ND_B200_KERNEL=jag,ND_B200_RK4_COOP=0;fused:ND_B200_KERNEL=fused"
timeout 600 python tools/bench_configs.py cfg4 cfg1 --check "--modes=$MODESR" > gpurun_out/r02g_sweep_rk4_coop.jsonl 2>> gpurun_out/r02g_sweep_persist.err
python tools/fmt_bench.py < gpurun_out/r02g_sweep_rk4_coop.jsonl
( time timeout 1500 python -m pytest tests/test_zzzz_full_size.py tests/test_gpu_parity.py -m gpu -x -q -k "full or rk4_trajectory or cfg" ) > gpurun_out/r02g_pytest_fullsize.log 2>&1
tail -n 15 gpurun_out/r02g_pytest_fullsize.log
