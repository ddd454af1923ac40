import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import run_model as rm
from golden_util import voxel_cases
from mdvcontainment_b200 import device as dev
name = sys.argv[1] if len(sys.argv) > 1 else "perlin_64x64x64_s3"
g = voxel_cases()[name]["grid"]
bits = dev.BitGrid.from_bool(g)
plan = dev.FramePlan(g.shape)
plan.label(bits); plan.pairs(bits)
h = plan.read_header()
rs, ro, lab, nt, nf = rm.label_runs(g)
R = h.total_runs
print("runs", R, len(rs) - 1, "labels", h.n_true, h.n_false, nt, nf)
d_ro = plan.run_table(0, g.shape[0] * g.shape[1] + 1)
d_rs = plan.run_table(1, R + 1)
d_lab = plan.run_table(3, R)
d_par = plan.run_table(2, R)
print("rowoff ok", np.array_equal(d_ro, ro), "runstart ok", np.array_equal(d_rs, rs))
bad = np.flatnonzero(d_lab != lab)
print("lab mismatches", len(bad), bad[:10], d_lab[bad[:10]], lab[bad[:10]])
if not np.array_equal(d_rs, rs):
    b = np.flatnonzero(d_rs != rs)
    print("first runstart mismatch", b[:10], d_rs[b[:10]], rs[b[:10]])
c, b = rm.pairs(g, rs, ro, lab)
print("contacts ok", np.array_equal(plan.contacts(), c), "bridges ok", np.array_equal(plan.bridges(), b), len(plan.bridges()), len(b))
