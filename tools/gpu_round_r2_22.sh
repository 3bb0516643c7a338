#!/bin/bash
# Round 2, GPU call 22: cooperative RK4 with one 1024-thread block per SM; ncu (full set + source) of the default kernels of
# configs 2 and 3; SASS of the benchmark kernels; coop / async-gather GPU tests
mkdir -p gpurun_out
MODESR="nocoop:;coop:ND_B200_RK4_COOP=1;jag_coop:ND_B200_KERNEL=jag,ND_B200_RK4_COOP=1;jag_nocoop:ND_B200_KERNEL=jag"
timeout 600 python tools/bench_configs.py cfg4 cfg1 --check "--modes=$MODESR" > gpurun_out/r02h_sweep_rk4_coop1024.jsonl 2> gpurun_out/r02h.err
python tools/fmt_bench.py < gpurun_out/r02h_sweep_rk4_coop1024.jsonl
( time timeout 900 python -m pytest tests/test_zzzz_round2_kernels.py -m gpu -x -q ) > gpurun_out/r02h_pytest_round2_kernels.log 2>&1
tail -n 6 gpurun_out/r02h_pytest_round2_kernels.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rhs_jag -s 8 -c 1 -f -o gpurun_out/r02h_jag_default_cfg2 python tools/bench_configs.py cfg2 --quick > gpurun_out/r02h_ncu_cfg2.log 2>&1
tail -n 2 gpurun_out/r02h_ncu_cfg2.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rhs_fused -s 8 -c 1 -f -o gpurun_out/r02h_fused_default_cfg3 python tools/bench_configs.py cfg3 --quick > gpurun_out/r02h_ncu_cfg3.log 2>&1
tail -n 2 gpurun_out/r02h_ncu_cfg3.log
