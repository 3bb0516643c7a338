#!/bin/bash
# Round 2, GPU call 23: round-2 kernel GPU tests; ncu (full set + source) of the default kernels of configs 2 and 3, summarised on
# the box (the reports with imported source exceed gpurun's 64 MiB return limit)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_zzzz_round2_kernels.py -m gpu -x -q ) > gpurun_out/r02h_pytest_round2_kernels.log 2>&1
tail -n 40 gpurun_out/r02h_pytest_round2_kernels.log
for c in cfg2:rhs_jag cfg3:rhs_fused; do
  cfg=${c%%:*}; k=${c#*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o /tmp/r02h_${cfg} python tools/bench_configs.py $cfg --quick > gpurun_out/r02h_ncu_${cfg}.log 2>&1
  tail -n 1 gpurun_out/r02h_ncu_${cfg}.log
  python tools/ncu_summary.py /tmp/r02h_${cfg}.ncu-rep > gpurun_out/r02h_${cfg}_ncu_summary.txt
  ncu -i /tmp/r02h_${cfg}.ncu-rep --page source --csv > gpurun_out/r02h_${cfg}_ncu_source.csv 2>/dev/null
  ncu -i /tmp/r02h_${cfg}.ncu-rep --page details > gpurun_out/r02h_${cfg}_ncu_details.txt 2>/dev/null
  ls -la /tmp/r02h_${cfg}.ncu-rep gpurun_out/r02h_${cfg}_ncu_source.csv
done
du -sh gpurun_out
