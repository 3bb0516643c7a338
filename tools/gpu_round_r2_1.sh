#!/bin/bash
# Round 2, GPU call 1 (one GPU): (1) raw gather-floor evidence: gather_bench + dsmem_bench; (2) the -m gpu suite at HEAD;
# (3) the prepared variants (packed edge parameters, degree-bucketed jagged windows) on cfg2 / cfg3.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_1_smi.txt 2>&1
timeout 300 tools/_build/gather_bench > gpurun_out/r02_1_gather_bench.txt 2>&1; echo "gather_bench rc=$?"
timeout 600 tools/_build/dsmem_bench > gpurun_out/r02_1_dsmem_bench.txt 2>&1; echo "dsmem_bench rc=$?"
tail -60 gpurun_out/r02_1_dsmem_bench.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_1_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_1_pytest_gpu.log
tail -8 gpurun_out/r02_1_pytest_gpu.log
MODES="default:;fused_packed:ND_B200_KERNEL=fused,ND_B200_PACK_P=1;jag128:ND_B200_KERNEL=jag,ND_B200_JAG_WINDOW=128;jag128_packed:ND_B200_KERNEL=jag,ND_B200_JAG_WINDOW=128,ND_B200_PACK_P=1;jag64_packed:ND_B200_KERNEL=jag,ND_B200_JAG_WINDOW=64,ND_B200_PACK_P=1"
timeout 900 python tools/bench_configs.py cfg2 cfg3 cfg2kura --check "--modes=$MODES" > gpurun_out/r02_1_sweep.jsonl 2> gpurun_out/r02_1_sweep.err
python tools/fmt_bench.py < gpurun_out/r02_1_sweep.jsonl
