#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python tools/e2e_sweep.py > gpurun_out/e2e_sweep.jsonl 2> gpurun_out/e2e_sweep.err
cat gpurun_out/e2e_sweep.jsonl
timeout 600 python tools/bench_configs.py cfg1 cfg2 cfg2nop cfg3 cfg4 --check > gpurun_out/sweep_c.jsonl 2> gpurun_out/sweep_c.err
python tools/fmt_bench.py < gpurun_out/sweep_c.jsonl
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json
timeout 600 python bench.py --steps 100 --warmup 5 --workload grid_kuramoto_1e6 --no-cpu-baseline > gpurun_out/bench_n1_grid.json 2> gpurun_out/bench_n1_grid.err
timeout 900 python bench.py --steps 50 --warmup 5 --workload cfg5_kuramoto_er_5e6 --no-cpu-baseline > gpurun_out/bench_n1_cfg5s.json 2> gpurun_out/bench_n1_cfg5s.err
tail -c 600 gpurun_out/bench_n1_grid.json; tail -c 600 gpurun_out/bench_n1_cfg5s.json
