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
    import ctypes
    from mdvcontainment_b200 import _lib
    lib = _lib.load()
    out, tmp = dev.BitGrid((n, n, n)), dev.BitGrid((n, n, n))
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for ops in ("de", "d", "e", "dde"):
        def run():
            rc = lib.mdvc_morph(bits0.words.data_ptr(), out.words.data_ptr(), tmp.words.data_ptr(), n, n, n, bits0.pitch,
                                ops.encode(), st)
            assert rc == 0
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(reps):
            run()
        ev[1].record()
        torch.cuda.synchronize()
        t = ev[0].elapsed_time(ev[1]) / reps
        launches = (len(ops) + 1) // 2
        print(f"morph '{ops}' {n}^3: {t * 1e3:8.1f} us   {launches * n ** 3 / 4 / t / 1e6:8.1f} GB/s algorithmic "
              f"({launches} launch(es), each reads + writes the bit grid once; preallocated buffers)")


if __name__ == "__main__":
    main()
