"""HAWK_TRACE of the N1 seam leg (second call)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

bench.variant_records_leg()
os.environ["HAWK_TRACE"] = "1"
print("==== traced", file=sys.stderr)
bench.variant_records_leg()
