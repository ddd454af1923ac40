"""Label / pairs stage times on three kinds of grids (development aid): smooth cylinder, blobs + salt noise, gyroid."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdvcontainment_b200 import device as dev, synth  # noqa: E402


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    ring = torch.from_numpy(synth.hollow_cylinder_grid(n, 1)[:, :, 0]).cuda()
    cases = {
        "cylinder": dev.BitGrid.from_bool(ring[:, :, None].expand(n, n, n).contiguous()),
        "salt2%": dev.BitGrid.from_bool(torch.from_numpy(synth.salt_noise_grid((n, n, n), salt=0.02)).cuda()),
        "gyroid": synth.gyroid_salt_planes(0, n, (n, n, n), salt=1e-4),
    }
    for name, bits in cases.items():
        plan, hdr = dev.label_with_retry(bits)
        tl = timed(lambda: plan.label(bits))
        tp = timed(lambda: plan.pairs(bits))
        print(f"{name:9s} {n}^3: label {tl:7.3f} ms  pairs {tp:7.3f} ms  runs={hdr.total_runs} +{hdr.n_true}/-{hdr.n_false}")


if __name__ == "__main__":
    main()
