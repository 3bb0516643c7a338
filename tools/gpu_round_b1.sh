#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python tools/e2e_sweep.py > gpurun_out/e2e_sweep.jsonl 2> gpurun_out/e2e_sweep.err
cat gpurun_out/e2e_sweep.jsonl
MODES="auto:;fused:ND_B200_KERNEL=fused;jag:ND_B200_KERNEL=jag"
timeout 600 python tools/bench_configs.py cfg1 cfg2 cfg3 cfg4 --check "--modes=$MODES" > gpurun_out/sweep_b.jsonl 2> gpurun_out/sweep_b.err
python tools/fmt_bench.py < gpurun_out/sweep_b.jsonl
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rhs_jag -s 8 -c 1 -f -o gpurun_out/jag_cfg4 python tools/bench_configs.py cfg4 --quick > gpurun_out/ncu_full_cfg4.log 2>&1
ls -la gpurun_out
