#!/bin/bash
# Round 2, GPU call 28: window mode of the jagged kernel (block = 128-row window: coalesced own outputs / du / vertex data)
mkdir -p gpurun_out
MODES="win:;nowin:ND_B200_JAG_WIN=0;win_pf:ND_B200_PF_DIST=600000;win_w48:ND_B200_JAG_WPS=48;win_w48_pf:ND_B200_JAG_WPS=48,ND_B200_PF_DIST=600000;nowin_pf:ND_B200_JAG_WIN=0,ND_B200_PF_DIST=600000"
timeout 900 python tools/bench_configs.py cfg2 cfg2kura cfg2nop cfg5s --check "--modes=$MODES" > gpurun_out/r02m_sweep_window_mode.jsonl 2> gpurun_out/r02m.err
python tools/fmt_bench.py < gpurun_out/r02m_sweep_window_mode.jsonl
tail -n 5 gpurun_out/r02m.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rhs_jag -s 8 -c 1 -f -o /tmp/r02m python tools/bench_configs.py cfg2 --quick > gpurun_out/r02m_ncu.log 2>&1
tail -n 1 gpurun_out/r02m_ncu.log
python tools/ncu_summary.py /tmp/r02m.ncu-rep > gpurun_out/r02m_jag_win_cfg2_ncu_summary.txt
ncu -i /tmp/r02m.ncu-rep --page source --csv > gpurun_out/r02m_jag_win_cfg2_ncu_source.csv 2>/dev/null
