#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_configs.py cfg1 cfg2 cfg2nop cfg3 cfg4 --check > gpurun_out/sweep_d.jsonl 2> gpurun_out/sweep_d.err
python tools/fmt_bench.py < gpurun_out/sweep_d.jsonl
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json
timeout 600 python bench.py --steps 100 --warmup 5 --workload grid_kuramoto_1e6 --no-cpu-baseline > gpurun_out/bench_n1_grid.json 2> gpurun_out/bench_n1_grid.err
tail -c 300 gpurun_out/bench_n1_grid.json
