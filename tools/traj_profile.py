import cProfile, pstats, sys, time, io
sys.path.insert(0, '/root/repo')
import torch
from mdvcontainment_b200 import synth
from mdvcontainment_b200.trajectory import PipelinePool, run_trajectory
n = 64
pinned = {}
for f in range(n):
    pos, dims = synth.trajectory_frame(f, box_nm=256.0)
    pinned[f] = torch.from_numpy(pos).pin_memory()
pool = PipelinePool(0.5, True, "", 3)
run = lambda: run_trajectory(lambda f: pinned[f], dims, 0.5, closing=True, n_frames=n, bind_numa=False, pool=pool)
run(); run()
def forget():
    for ps in pool.by_shape.values():
        for p in ps:
            if p is not None:
                p._solved.clear(); p._lut_key = None
ts = []
for _ in range(3):
    forget(); torch.cuda.synchronize(); t0 = time.perf_counter(); run(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("frames/s", [round(n / t, 1) for t in ts])
forget()
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45); print(s.getvalue())
