"""Time the closing ('de') alone on the 1024^3 vesicle scene -- development aid.

    python tools/morph_only.py [n] [reps]
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
    _, nbox, T, invT = voxel_transform(dims, 0.5)
    bits0 = dev.BitGrid((n, n, n))
    dev.bin_atoms(torch.from_numpy(pos).cuda(), T, invT, nbox, bits0)
    for _ in range(3):
        dev.morph(bits0, "de")
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        dev.morph(bits0, "de")
    ev[1].record()
    torch.cuda.synchronize()
    t = ev[0].elapsed_time(ev[1]) / reps
    print(f"close {n}^3: {t * 1e3:8.1f} us   {n ** 3 / 4 / t / 1e6:8.1f} GB/s algorithmic (read+write of the bit grid)")


if __name__ == "__main__":
    main()
