#!/bin/bash
# Round 2, GPU call 12: the -m gpu suite with the new defaults (auto window-128 jagged layout, compact entry words, adaptive
# packing), the config sweep with defaults, the bench line
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_12_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_12_pytest_gpu.log
tail -n 12 gpurun_out/r02_12_pytest_gpu.log
timeout 900 python tools/bench_configs.py cfg1 cfg2 cfg2nop cfg2kura cfg3 cfg4 cfg5s --check "--modes=auto:;auto_pk:ND_B200_PACK_P=1" > gpurun_out/r02_12_sweep_defaults.jsonl 2> gpurun_out/r02_12_sweep_defaults.err
python tools/fmt_bench.py < gpurun_out/r02_12_sweep_defaults.jsonl
tail -n 5 gpurun_out/r02_12_sweep_defaults.err
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r02_12_bench.json 2> gpurun_out/r02_12_bench.err
cat gpurun_out/r02_12_bench.json; tail -n 5 gpurun_out/r02_12_bench.err
