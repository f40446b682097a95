"""N1 through the seam on config 2's region (200 samples): where the host time goes."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

r = bench.variant_records_leg()
print({k: v for k, v in r.items() if k != "what"})
pr = cProfile.Profile()
pr.enable()
bench.variant_records_leg()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
