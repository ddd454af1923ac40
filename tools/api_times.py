"""Wall-clock of the drop-in API calls on the config-3 scene (development aid): python tools/api_times.py [n]"""
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdvcontainment_b200 import Containment, synth  # noqa: E402
from mdvcontainment_b200.universe import Universe  # noqa: E402


def timed(label, fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    print(f"{label:58s} {(time.perf_counter() - t) * 1e3:9.1f} ms")
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    pos, dims = synth.vesicle_scene(box_nm=n / 2.0)
    def quiet(f):
        def run():
            with contextlib.redirect_stdout(io.StringIO()):
                return f()
        return run
    for rep in range(2):
        u = Universe(pos.copy(), dims)
        timed(f"[{rep}] Containment(all, 0.5, closing=True, no_mapping=True, betafactors=False)",
              quiet(lambda: Containment(u.atoms, 0.5, closing=True, no_mapping=True, betafactors=False)))
        u = Universe(pos.copy(), dims)
        c = timed(f"[{rep}] Containment(all, 0.5, closing=True)   # mapping + tempfactors",
                  quiet(lambda: Containment(u.atoms, 0.5, closing=True)))
        timed(f"[{rep}] c.get_atomgroup_from_nodes([1], containment=True)", lambda: c.get_atomgroup_from_nodes([1], containment=True))
        timed(f"[{rep}] c.voxel_containment.components_grid (D2H of the int32 grid)", lambda: c.voxel_containment.components_grid)
        timed(f"[{rep}] str(c)", lambda: str(c))
        grid = timed(f"[{rep}] c.boolean_grid (1 GB bool to NumPy)", lambda: c.boolean_grid)
        from mdvcontainment_b200 import VoxelContainment
        timed(f"[{rep}] VoxelContainment(numpy bool grid)", quiet(lambda: VoxelContainment(grid)))
        del c
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
