#!/bin/bash
# Round-2 ncu evidence (run under gpurun, one GPU): launch lists of the bench command for configs 2
# and 4, and --set full captures of the kernels this round added.
set -x
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-clocks > $O/r02_ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-clocks > $O/r02_ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fused_scan|resolve_count|resolve_write|fused_expand" -s 4 -c 4 -o $O/r02_c4_kernels python tools/one_step.py c4 > $O/r02_ncu_c4k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fused_scan" -s 1 -c 1 -o $O/r02_c2_fused python tools/one_step.py c2 1.0 fused > $O/r02_ncu_c2f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pack_kernel|match_kernel|rows_fast|gather_fast" -s 4 -c 4 -o $O/r02_c2_kernels python tools/one_step.py c2 1.0 staged > $O/r02_ncu_c2k.log 2>&1
tail -2 $O/r02_ncu_c2k.log
