"""Throughput of run_trajectory on pre-generated frames of the config-4 scene (development aid):
python tools/traj_bench.py [box_nm=256] [frames=40]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdvcontainment_b200 import synth  # noqa: E402
from mdvcontainment_b200.trajectory import run_trajectory  # noqa: E402

box = float(sys.argv[1]) if len(sys.argv) > 1 else 256.0
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 40
frames = [synth.trajectory_frame(f, box_nm=box) for f in range(8)]
dims = frames[0][1]
pos = [frames[f % 8][0] for f in range(nf)]
run_trajectory(pos[:6], dims, 0.5)                      # warm-up: builds, capacities, graphs
torch.cuda.synchronize()
t = time.perf_counter()
res = run_trajectory(pos, dims, 0.5)
torch.cuda.synchronize()
dt = time.perf_counter() - t
print(f"{nf} frames of {int(box * 2)}^3 ({len(pos[0])} beads): {nf / dt:.1f} frames/s wall-clock (positions from pageable NumPy)")
print(res[0]["str"].splitlines()[0])
