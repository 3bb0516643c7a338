#!/bin/bash
# Round 2, GPU call 46: L2 fetch granularity hint on a DRAM-resident gather table (Kuramoto on ER 2e7 / 1.6e8: u = 160 MB > L2)
mkdir -p gpurun_out
ND_PROFILE_L2_GRANULARITY=1 timeout 600 python tools/profile_cfg5_full.py 20000000 160000000 > gpurun_out/r02ab_l2_granularity.log 2>&1
cat gpurun_out/r02ab_l2_granularity.log | tail -8
