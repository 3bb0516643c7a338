#!/bin/bash
# Round 2, GPU call 33 (2 GPUs): compute-sanitizer memcheck / racecheck / synccheck on smoke() and on the 2-GPU p2p parity tests
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  ( time timeout 600 $CS --tool $tool --target-processes all --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02q_sanitizer_${tool}_smoke.log 2>&1
  tail -n 6 gpurun_out/r02q_sanitizer_${tool}_smoke.log
done
for tool in memcheck racecheck; do
  ( time timeout 900 $CS --tool $tool --target-processes all --print-limit 20 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "p2p" ) > gpurun_out/r02q_sanitizer_${tool}_two_gpu_p2p.log 2>&1
  tail -n 12 gpurun_out/r02q_sanitizer_${tool}_two_gpu_p2p.log
done
# the round-2 single-GPU kernels (jagged default with L2 prefetch, window mode, batched, async) under memcheck at small size
( time timeout 900 $CS --tool memcheck --target-processes all --print-limit 20 python -m pytest tests/test_zzzz_round2_kernels.py -m gpu -x -q -k "window_mode or chunked" ) > gpurun_out/r02q_sanitizer_memcheck_round2_kernels.log 2>&1
tail -n 8 gpurun_out/r02q_sanitizer_memcheck_round2_kernels.log
