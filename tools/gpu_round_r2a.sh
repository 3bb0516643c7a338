#!/bin/bash
# Round 2, first GPU call (one GPU, ~25 min): everything written after round 1's GPU budget ended runs on the B200 for the
# first time here.  (1) the whole -m gpu suite (stateful edges, backend-parametrised tests, ND_LAUNCH build);
# (2) the degree-bucketed jagged layout (ND_B200_JAG_WINDOW) against the default kernels on every single-GPU config;
# (3) the contract bench line; (4) ncu launch list + one full capture of whichever jagged window wins on cfg2.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
MODES="default:;fused:ND_B200_KERNEL=fused;jag32:ND_B200_KERNEL=jag;jag64:ND_B200_KERNEL=jag,ND_B200_JAG_WINDOW=64;jag128:ND_B200_KERNEL=jag,ND_B200_JAG_WINDOW=128;jag128_u4:ND_B200_KERNEL=jag,ND_B200_JAG_WINDOW=128,ND_B200_JAG_U=4;jag128_w64:ND_B200_KERNEL=jag,ND_B200_JAG_WINDOW=128,ND_B200_JAG_WPS=64;packed:ND_B200_PACK_P=1;fused_packed:ND_B200_KERNEL=fused,ND_B200_PACK_P=1;jag128_packed:ND_B200_KERNEL=jag,ND_B200_JAG_WINDOW=128,ND_B200_PACK_P=1;rk4pack:ND_B200_RK4_PACK=1"
timeout 1200 python tools/bench_configs.py cfg1 cfg2 cfg2nop cfg2kura cfg3 cfg4 cfg5s --check "--modes=$MODES" > gpurun_out/r02a_sweep_window.jsonl 2> gpurun_out/r02a_sweep_window.err
python tools/fmt_bench.py < gpurun_out/r02a_sweep_window.jsonl
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/r02a_bench_cfg2.json 2> gpurun_out/r02a_bench_cfg2.err
cat gpurun_out/r02a_bench_cfg2.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02a_bench_cfg2_reference_arm.json 2>> gpurun_out/r02a_bench_cfg2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ND_B200_KERNEL=jag ND_B200_JAG_WINDOW=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rhs_jag -s 8 -c 1 -f -o gpurun_out/r02a_jag128_cfg2 python tools/bench_configs.py cfg2 --quick > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
