#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/js_check.py 200000 > gpurun_out/r02_5_js_check.txt 2>&1
tail -30 gpurun_out/r02_5_js_check.txt
ND_B200_KERNEL=js timeout 400 ncu --set full --clock-control none --import-source on -k regex:rhs_js -s 8 -c 1 -f -o gpurun_out/r02_5_js88_pk_cfg2 python tools/bench_configs.py cfg2 --quick "--modes=jp:ND_B200_KERNEL=js,ND_B200_PACK_P=1" > gpurun_out/r02_5_ncu.log 2>&1
tail -n 3 gpurun_out/r02_5_ncu.log
ND_B200_KERNEL=js ND_B200_JS_U=4 ND_B200_PACK_P=1 timeout 600 compute-sanitizer --tool racecheck python tools/js_check.py 20000 > gpurun_out/r02_5_racecheck.txt 2>&1
tail -n 15 gpurun_out/r02_5_racecheck.txt
