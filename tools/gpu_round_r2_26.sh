#!/bin/bash
# Round 2, GPU call 26: ncu (full set + source) of rhs_jagb_kernel CH=8 / 32 warps on config 2, summarised on the box
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rhs_jagb -s 8 -c 1 -f -o /tmp/r02k python tools/bench_configs.py cfg2 --quick "--modes=b:ND_B200_KERNEL=jagb,ND_B200_JAG_WINDOW=128,ND_B200_JAGA_WPS=32" > gpurun_out/r02k_ncu.log 2>&1
tail -n 1 gpurun_out/r02k_ncu.log
python tools/ncu_summary.py /tmp/r02k.ncu-rep > gpurun_out/r02k_jagb8w32_cfg2_ncu_summary.txt
ncu -i /tmp/r02k.ncu-rep --page source --csv > gpurun_out/r02k_jagb8w32_cfg2_ncu_source.csv 2>/dev/null
ncu -i /tmp/r02k.ncu-rep --page details > gpurun_out/r02k_jagb8w32_cfg2_ncu_details.txt 2>/dev/null
