set -x
O=gpurun_out
python bench.py --workload c4 --steps 20 --warmup 3 > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-clocks > $O/r02_ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fused_scan|resolve_count|resolve_write|fused_expand" -s 4 -c 4 -o $O/r02_c4_kernels python tools/one_step.py c4 > $O/r02_ncu_c4k.log 2>&1
tail -1 $O/r02_ncu_c4k.log
