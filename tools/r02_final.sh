#!/bin/bash
# Final round-2 lines (one GPU): tests, then the bench lines the profiles/ directory keeps
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
tail -2 $O/r02_bench_c5shard.err
