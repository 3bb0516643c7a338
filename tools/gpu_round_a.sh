#!/bin/bash
# one GPU call: parity tests, mode sweep on every single-GPU config, the bench line, ncu launch list + one full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
MODES="fused:ND_B200_KERNEL=fused;jag_u2_w32:ND_B200_KERNEL=jag,ND_B200_JAG_U=2,ND_B200_JAG_WPS=32;jag_u2_w48:ND_B200_KERNEL=jag,ND_B200_JAG_U=2,ND_B200_JAG_WPS=48;jag_u2_w64:ND_B200_KERNEL=jag,ND_B200_JAG_U=2,ND_B200_JAG_WPS=64;jag_u4_w32:ND_B200_KERNEL=jag,ND_B200_JAG_U=4,ND_B200_JAG_WPS=32;jag_u4_w48:ND_B200_KERNEL=jag,ND_B200_JAG_U=4,ND_B200_JAG_WPS=48;jag_u4_w64:ND_B200_KERNEL=jag,ND_B200_JAG_U=4,ND_B200_JAG_WPS=64"
timeout 900 python tools/bench_configs.py cfg1 cfg2 cfg2nop cfg3 cfg4 --check "--modes=$MODES" > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
python tools/fmt_bench.py < gpurun_out/sweep.jsonl
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
ND_B200_KERNEL=fused timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err
cat gpurun_out/bench_default.json
# ncu: launch list of the bench command, then one full capture of the dominant kernel (never a bench value)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rhs_jag -s 8 -c 1 -f -o gpurun_out/jag_cfg2 python tools/bench_configs.py cfg2 --quick > gpurun_out/ncu_full.log 2>&1
ND_B200_KERNEL=fused timeout 600 ncu --set full --clock-control none --import-source on -k regex:rhs_fused -s 8 -c 1 -f -o gpurun_out/fused_cfg2 python tools/bench_configs.py cfg2 --quick > gpurun_out/ncu_full_fused.log 2>&1
ls -la gpurun_out
