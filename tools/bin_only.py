"""Time the binning kernel alone on the config-3 vesicle scene -- development aid.

    python tools/bin_only.py [n] [reps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdvcontainment_b200 import device as dev, synth  # noqa: E402
from mdvcontainment_b200.atomgroup_to_voxels import voxel_transform  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    pos, dims = synth.vesicle_scene(box_nm=n / 2.0)
    order = os.environ.get("BIN_ORDER", "none")          # locality experiment: none | slabs<k> | full
    if order != "none":
        import numpy as np
        vx = np.floor(pos[:, 0] / 5.0).astype(np.int64) % n
        if order == "full":
            vy = np.floor(pos[:, 1] / 5.0).astype(np.int64) % n
            key = vx * n + vy
        else:
            key = vx // (n // int(order[5:]))
        pos = pos[np.argsort(key, kind="stable")]
    _, nbox, T, invT = voxel_transform(dims, 0.5)
    bits0 = dev.BitGrid((n, n, n))
    p = torch.from_numpy(pos).cuda()
    thrash = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # evicts the L2 between repetitions
    scratch = dev.bin_scratch(len(pos), "cuda")

    def plain():
        bits0.words.zero_()
        dev.bin_atoms(p, T, invT, nbox, bits0)

    def bucketed():
        dev.bin_grid(p, T, invT, nbox, bits0, scratch)

    for name, fn in (("clear + mdvc_bin_atoms", plain), ("mdvc_bin_grid", bucketed)):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            thrash.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        t = ts[len(ts) // 2]
        print(f"{name:24s} {len(pos)} atoms into {n}^3: {t * 1e3:8.1f} us median (L2 evicted before each), "
              f"{len(pos) * 12 / t / 1e6:8.1f} GB/s (12 B/atom)")


if __name__ == "__main__":
    main()
