#!/bin/bash
# Round 2, GPU call 3: fresh ncu --set full captures from HEAD: default tile kernel on cfg2 and cfg3, packed jagged (window 128) on cfg2
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rhs_fused -s 8 -c 1 -f -o gpurun_out/r02_3_fused_cfg2 python tools/bench_configs.py cfg2 --quick > gpurun_out/r02_3_ncu_a.log 2>&1
ND_B200_KERNEL=jag ND_B200_JAG_WINDOW=128 timeout 400 ncu --set full --clock-control none --import-source on -k regex:rhs_jag -s 8 -c 1 -f -o gpurun_out/r02_3_jag128_packed_cfg2 python tools/bench_configs.py cfg2 --quick "--modes=jp:ND_B200_KERNEL=jag,ND_B200_JAG_WINDOW=128,ND_B200_PACK_P=1" > gpurun_out/r02_3_ncu_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rhs_fused -s 8 -c 1 -f -o gpurun_out/r02_3_fused_cfg3 python tools/bench_configs.py cfg3 --quick > gpurun_out/r02_3_ncu_c.log 2>&1
tail -3 gpurun_out/r02_3_ncu_*.log
ls -la gpurun_out
