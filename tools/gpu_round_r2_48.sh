#!/bin/bash
# Round 2, GPU call 48: column-blocked evaluation (ND_B200_L2_BLOCKS): GPU leg of its parity test, then single pass vs blocks on a
# graph whose vertex outputs exceed the L2 (Kuramoto on ER 2e7 / 1.6e8)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_zzzz_round2_kernels.py -m gpu -x -q -k "column_blocked" ) > gpurun_out/r02ad_pytest_blocks.log 2>&1
tail -n 4 gpurun_out/r02ad_pytest_blocks.log
timeout 900 python tools/profile_l2_blocks.py 20000000 160000000 auto 8 > gpurun_out/r02ad_l2_blocks.log 2>&1
cat gpurun_out/r02ad_l2_blocks.log | tail -6
