#!/bin/bash
# Final lines after the K1 rewrite (pack_chunk_v3, cp.async-fed fused kernel as the default): tests,
# the bench lines profiles/ keeps, the ncu launch list of the bench command and --set full captures
# of the kernels of a config-2 step (one GPU; bench numbers never come from a run under ncu)
set -x
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_c2_reference_arm.json 2> $O/r02_ref_c2.err
python bench.py --steps 20 --warmup 3 > $O/r02_bench_c2.json 2> $O/r02_bench_c2.err
python bench.py --workload c3 --steps 20 --warmup 3 > $O/r02_bench_c3.json 2> $O/r02_bench_c3.err
python bench.py --workload c4 --steps 20 --warmup 3 > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err
python bench.py --impl reference --workload c4 --steps 2 --warmup 1 > $O/r02_bench_c4_reference_arm.json 2> $O/r02_ref_c4.err
python bench.py --workload c5shard --steps 10 --warmup 3 --no-e2e --no-cpu > $O/r02_bench_c5shard.json 2> $O/r02_bench_c5shard.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-clocks > $O/r02_ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fused_scan|fused_expand|rows_fast|gather_fast" -s 4 -c 4 -o $O/r02_c2_final_kernels python tools/one_step.py c2 > $O/r02_ncu_c2k.log 2>&1
tail -2 $O/r02_ncu_c2k.log
for f in c2 c3 c4 c5shard; do tail -c 300 $O/r02_bench_$f.err; done
