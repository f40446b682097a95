#!/bin/bash
# Round-2 evidence (run under gpurun, one GPU). Bench lines first (never under ncu), then the ncu
# launch lists of the same commands, then --set full captures of the kernels this round added.
set -x
O=gpurun_out
python bench.py --steps 20 --warmup 3 > $O/r02_bench_c2.json 2> $O/r02_bench_c2.err
python bench.py --workload c3 --steps 20 --warmup 3 > $O/r02_bench_c3.json 2> $O/r02_bench_c3.err
python bench.py --workload c4 --steps 20 --warmup 3 > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_c2_reference_arm.json 2> $O/r02_ref_c2.err
python bench.py --impl reference --workload c4 --steps 2 --warmup 1 > $O/r02_bench_c4_reference_arm.json 2> $O/r02_ref_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-clocks > $O/r02_ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-clocks > $O/r02_ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fused_scan|resolve_count|resolve_write|fused_expand" -s 4 -c 4 -o $O/r02_c4_kernels python tools/one_step.py c4 > $O/r02_ncu_c4k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"edit_windows|edits_plain|derive_kernel" -s 3 -c 3 -o $O/r02_edit_planes python tools/edits_trace.py > $O/r02_ncu_edits.log 2>&1
tail -2 $O/r02_ncu_edits.log
