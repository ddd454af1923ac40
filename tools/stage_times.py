"""Per-stage device timings of the frame pipeline (CUDA events) -- development aid, not the bench.

    python tools/stage_times.py [n] [reps]      hollow cylinder n^3 (default 512), closing 'de'
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdvcontainment_b200 import device as dev, synth  # noqa: E402


def timed(fn, reps):
    torch.cuda.synchronize()
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    kind = sys.argv[3] if len(sys.argv) > 3 else "cylinder"
    if kind == "cylinder":
        ring = torch.from_numpy(synth.hollow_cylinder_grid(n, 1)[:, :, 0]).cuda()
        g = ring[:, :, None].expand(n, n, n).contiguous()
    else:
        pos, dims = synth.vesicle_scene(box_nm=n / 2.0)
        print(f"atoms {len(pos)}")
        g = None
    shape = (n, n, n)
    N = n ** 3
    if g is not None:
        bits0 = dev.BitGrid.from_bool(g)
        del g
    else:
        from mdvcontainment_b200.atomgroup_to_voxels import voxel_transform
        _, nbox, T, invT = voxel_transform(dims, 0.5)
        p = torch.from_numpy(pos).cuda()
        bits0 = dev.BitGrid(shape)
        dev.bin_atoms(p, T, invT, nbox, bits0)
        t = timed(lambda: dev.bin_atoms(p, T, invT, nbox, bits0), reps)
        print(f"bin_atoms        {t:9.3f} ms   {len(pos) * 12 / t / 1e6:8.1f} GB/s (12 B/atom)")
    t = timed(lambda: dev.morph(bits0, "de"), reps)
    print(f"morph 'de'       {t:9.3f} ms   {N * 0.25 / t / 1e6:8.1f} GB/s (0.25 B/voxel)")
    bits = dev.morph(bits0, "de")
    plan = dev.FramePlan(shape)
    print(f"workspace {plan.nbytes / 2**20:.1f} MiB, caps runs={plan.caps.max_runs} labels={plan.caps.max_labels}")
    t = timed(lambda: plan.label(bits), reps)
    print(f"frame_label      {t:9.3f} ms")
    t = timed(lambda: plan.pairs(bits), reps)
    print(f"frame_pairs      {t:9.3f} ms")
    h = plan.read_header()
    print(f"runs={h.total_runs} +{h.n_true}/-{h.n_false} contacts={h.n_contacts} bridges={h.n_bridges} flags={h.flags}")
    t = timed(lambda: plan.components(True), reps)
    print(f"frame_components {t:9.3f} ms")
    labels = torch.empty(shape, dtype=torch.int32, device="cuda")
    comps = torch.empty(shape, dtype=torch.int32, device="cuda")
    t = timed(lambda: plan.expand(bits, labels, None), reps)
    print(f"expand labels    {t:9.3f} ms   {N * 4 / t / 1e6:8.1f} GB/s (4 B/voxel)")
    t = timed(lambda: plan.expand(bits, labels, comps), reps)
    print(f"expand both      {t:9.3f} ms   {N * 8 / t / 1e6:8.1f} GB/s (8 B/voxel)")

    def frame():
        b = dev.morph(bits0, "de")
        plan.label(b); plan.pairs(b); plan.components(True); plan.expand(b, labels, comps)
    t = timed(frame, reps)
    print(f"whole frame      {t:9.3f} ms   {N * 8.5 / t / 1e6:8.1f} GB/s (8.5 B/voxel)  {1e3 / t:7.1f} frames/s")
    h = plan.read_header()
    print("counts", plan.component_counts().tolist())
    a = torch.empty(1 << 28, dtype=torch.int32, device="cuda"); b = torch.empty_like(a)
    t = timed(lambda: b.copy_(a), reps)
    print(f"torch copy 1 GiB {t:9.3f} ms   {2 * a.numel() * 4 / t / 1e6:8.1f} GB/s")
    t = timed(lambda: a.fill_(1), reps)
    print(f"torch fill 1 GiB {t:9.3f} ms   {a.numel() * 4 / t / 1e6:8.1f} GB/s")


if __name__ == "__main__":
    main()
