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
    for _ in range(3):
        dev.bin_atoms(p, T, invT, nbox, bits0)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        dev.bin_atoms(p, T, invT, nbox, bits0)
    ev[1].record()
    torch.cuda.synchronize()
    t = ev[0].elapsed_time(ev[1]) / reps
    print(f"bin {len(pos)} atoms into {n}^3: {t * 1e3:8.1f} us   {len(pos) * 12 / t / 1e6:8.1f} GB/s (12 B/atom)")


if __name__ == "__main__":
    main()
