#!/bin/bash
# Round 2, GPU call 49: -m gpu at HEAD (after the column-block plumbing)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02ae_pytest_gpu.log 2>&1
tail -n 5 gpurun_out/r02ae_pytest_gpu.log
