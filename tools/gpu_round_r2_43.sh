#!/bin/bash
# Round 2, GPU call 43: the driver's round-end sequence at HEAD -- build(), -m gpu, smoke()
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.build(); g.smoke()" ) > gpurun_out/r02y_build_smoke.log 2>&1
tail -n 5 gpurun_out/r02y_build_smoke.log
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02y_pytest_gpu.log 2>&1
tail -n 6 gpurun_out/r02y_pytest_gpu.log
