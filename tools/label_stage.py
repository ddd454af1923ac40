"""Run the label + pairs stages of the 1024^3 vesicle scene a few times (driver for ncu captures of the run-level
kernels):   python tools/label_stage.py [n=1024] [reps=3]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdvcontainment_b200 import device as dev, synth  # noqa: E402
from mdvcontainment_b200.atomgroup_to_voxels import voxel_transform  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    pos, dims = synth.vesicle_scene(box_nm=n / 2.0)
    _, nbox, T, invT = voxel_transform(dims, 0.5)
    bits0 = dev.BitGrid((n, n, n))
    dev.bin_grid(torch.from_numpy(pos).cuda(), T, invT, nbox, bits0)
    closed = dev.morph(bits0, "de")
    if "--empty" in sys.argv:                       # one run per row: the dense write with nothing to look up
        closed.words.zero_()
    plan = dev.FramePlan((n, n, n))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for r in range(reps):
        ev[0].record()
        plan.label(closed)
        ev[1].record()
        plan.pairs(closed)
        ev[2].record()
        torch.cuda.synchronize()
        h = plan.read_header()
        print(f"[{r}] label {ev[0].elapsed_time(ev[1]) * 1e3:7.1f} us  pairs {ev[1].elapsed_time(ev[2]) * 1e3:7.1f} us  "
              f"runs {h.total_runs} labels +{h.n_true}/-{h.n_false} contacts {h.n_contacts} bridges {h.n_bridges} flags {h.flags} "
              f"pair rows {int(plan.ws[:256].view(torch.int32)[11])}")
    # the dense write alone (components from the device-side periodic merge), labels + components and components only
    plan.components(True)
    labels = torch.empty((n, n, n), dtype=torch.int32, device="cuda")
    comps = torch.empty((n, n, n), dtype=torch.int32, device="cuda")
    for name, args in (("labels + components", (labels, comps)), ("components only", (None, comps))):
        ts = []
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.expand(closed, *args)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        nbytes = n ** 3 * (4.0 * sum(a is not None for a in args) + 0.125)
        print(f"expand {name:20s} {t * 1e3:8.1f} us  {nbytes / t / 1e6:7.0f} GB/s   digest {int(comps[::97, ::89, ::83].sum())}")


if __name__ == "__main__":
    main()
