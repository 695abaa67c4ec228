"""Short dense-batch run for ncu: S=1024, A=4, B=1024 candidates, a few soft-VI sweeps (csrc/dense_batch.cu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))
import bench
print(bench.dense_batch_line(B=1024, sweeps=6, B_gather=32))
