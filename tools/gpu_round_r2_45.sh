#!/bin/bash
# Round 2, GPU call 45: ncu (full set) of the default kernel on the full configs[4] graph at N=1 (gathers from DRAM)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:rhs_jag -s 7 -c 1 -f -o /tmp/r02aa python tools/profile_cfg5_full.py > gpurun_out/r02aa_ncu_cfg5_full.log 2>&1
tail -n 3 gpurun_out/r02aa_ncu_cfg5_full.log
python tools/ncu_summary.py /tmp/r02aa.ncu-rep > gpurun_out/r02aa_jag_cfg5_full_ncu_summary.txt
ncu -i /tmp/r02aa.ncu-rep --page details > gpurun_out/r02aa_jag_cfg5_full_ncu_details.txt 2>/dev/null
grep -E "gpu__time_duration|dram__bytes|lts__t_sectors.sum|t_sectors_pipe_lsu_mem_global_op_ld.sum|lts__t_sector_hit" gpurun_out/r02aa_jag_cfg5_full_ncu_summary.txt
