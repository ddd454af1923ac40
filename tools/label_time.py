"""Time mdvc_frame_label / mdvc_frame_pairs on the vesicle scene (development aid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdvcontainment_b200 import device as dev, synth
from mdvcontainment_b200.pipeline import FramePipeline

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
pos, dims = synth.vesicle_scene(box_nm=size / 2.0, seed=7)
pipe = FramePipeline(dims, 0.5, closing=True, write_labels=False)
bits = pipe.voxelize(torch.from_numpy(pos).cuda())

def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

tl = timed(lambda: pipe.plan.label(bits))
tp = timed(lambda: pipe.plan.pairs(bits))
h = pipe.plan.read_header()
print(f"RY={os.environ.get('MDVC_LINK_RADIX_Y','-')} RX={os.environ.get('MDVC_LINK_RADIX_X','-')} FL={os.environ.get('MDVC_LINK_FLATTEN','-')}: "
      f"label {tl:.3f} ms  pairs {tp:.3f} ms  runs={h.total_runs} +{h.n_true}/-{h.n_false} c={h.n_contacts} b={h.n_bridges}")
