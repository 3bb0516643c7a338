#!/bin/bash
# Round 2, GPU call 47: L2 fetch-size hints on the gather loads, DRAM-resident gather table (Kuramoto on ER 2e7 / 1.6e8)
mkdir -p gpurun_out
ND_PROFILE_GATHER_HINT=1 timeout 600 python tools/profile_cfg5_full.py 20000000 160000000 > gpurun_out/r02ac_gather_hint.log 2>&1
tail -6 gpurun_out/r02ac_gather_hint.log
