import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from mdvcontainment_b200 import synth
from mdvcontainment_b200.pipeline import FramePipeline
pos, dims = synth.vesicle_scene(box_nm=512.0, seed=7)
pipe = FramePipeline(dims, 0.5, closing=True)
p = torch.from_numpy(pos).cuda()
for _ in range(3): out = pipe.run(p)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    t0 = time.perf_counter(); pipe.begin(p); t1 = time.perf_counter(); out = pipe.finish(); t2 = time.perf_counter()
    ts.append((t1 - t0, t2 - t1, out.timings))
torch.cuda.synchronize()
print("begin (enqueue) ms", np.mean([t[0] for t in ts]) * 1e3, "finish ms", np.mean([t[1] for t in ts]) * 1e3)
print({k: round(np.mean([t[2][k] for t in ts]) * 1e3, 3) for k in ts[0][2]})
