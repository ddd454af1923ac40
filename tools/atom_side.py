"""Exercise the atom-side kernels (A2/A9/N2: voxel-sorted atom list, node -> atoms selection, explicit voxel gather,
per-atom component lookup) on the config-3 scene with solvent, for timing and ncu captures (development aid):

    python tools/atom_side.py [n=1024]
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:'k_atom_keys|k_rs_|k_asel_|k_vox_|k_atom_run_values|k_key_run_values|k_bin_atoms4' --csv ... python tools/atom_side.py
"""
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
    print(f"{label:64s} {(time.perf_counter() - t) * 1e3:9.2f} ms")
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    pos, dims = synth.vesicle_scene(box_nm=n / 2.0)
    rng = np.random.default_rng(99)
    water = np.round(rng.random((len(pos), 3)) * dims[:3], 2).astype(np.float32)     # as many solvent beads again
    names = np.array(["C"] * len(pos) + ["W"] * len(water), dtype=object)
    u = Universe(np.concatenate([pos, water]), dims, names)
    sel = u.select_atoms("name C")
    print(f"{len(u.atoms)} atoms in the universe, {len(sel)} selected, grid {n}^3")
    for rep in range(2):
        with contextlib.redirect_stdout(io.StringIO()):
            c = timed(f"[{rep}] Containment(selection, 0.5, closing=True)  # mapping + tempfactors", lambda: Containment(sel, 0.5, closing=True))
        nodes = c.voxel_containment.nodes
        timed(f"[{rep}] voxel-sorted atom list (stable LSD radix sort, first use)", lambda: c.voxel2atom.csr)
        for node in nodes[:3]:
            ag = timed(f"[{rep}] get_atomgroup_from_nodes([{int(node)}], containment=True)", lambda: c.get_atomgroup_from_nodes([node], containment=True))
            print(f"        -> {len(ag)} atoms")
        vox = c.voxel_containment.get_voxel_positions([nodes[-1]])[:200000]
        ag = timed(f"[{rep}] get_atomgroup_from_voxel_positions({len(vox)} voxels)", lambda: c.get_atomgroup_from_voxel_positions(vox))
        print(f"        -> {len(ag)} atoms")
        buf = io.StringIO()

        def beta():
            with contextlib.redirect_stdout(buf):
                c.set_betafactors()
        timed(f"[{rep}] set_betafactors()", beta)
        del c
    print("done")


if __name__ == "__main__":
    main()
