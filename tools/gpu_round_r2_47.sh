#!/bin/bash
# Round 2, GPU call 47: L2 fetch-size hints on the gather loads, DRAM-resident gather table (Kuramoto on ER 2e7 / 1.6e8)
# (record of a finished experiment: the ND_B200_GATHER_HINT switch in the kernel and the ND_PROFILE_GATHER_HINT leg of the script were
#  removed again after this run -- no gain, profiles/r02ac_gather_l2_hint.log -- so this script no longer reproduces it)
mkdir -p gpurun_out
ND_PROFILE_GATHER_HINT=1 timeout 600 python tools/profile_cfg5_full.py 20000000 160000000 > gpurun_out/r02ac_gather_hint.log 2>&1
tail -6 gpurun_out/r02ac_gather_hint.log
