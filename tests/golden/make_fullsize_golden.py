"""Known answers at BASELINE.json's full size (config 3: ~10.2 M beads binned into 1024^3), produced WITHOUT the
CUDA path:

    oracle C binning (oracle.c: orc_bin_atoms)  ->  NumPy periodic closing (voxel_oracle.morph)
    ->  scipy.ndimage.label on both polarities, the reference's own labeller (label.py:7-17)
    ->  contacts / bridges by the reference's compiled Cython kernels (oracle/_ref) when they are present,
        else by the oracle's C restatement.

Whole grids are recorded as 64-bit digests (golden_util.tensor_digest), the small results verbatim.
Needs ~20 GB of host memory and a few minutes of one core; run once and commit the JSON:

    python tests/golden/make_fullsize_golden.py [n=1024]
"""
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(HERE)]

import voxel_oracle as vo  # noqa: E402
from golden_util import sha, tensor_digest  # noqa: E402
from mdvcontainment_b200 import synth  # noqa: E402  (input generation only: NumPy, no device code)


def digest(a):
    t = torch.from_numpy(a)
    if torch.cuda.is_available():
        return tensor_digest(t.cuda())
    return tensor_digest(t)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(HERE, f"fullsize_vesicle_{n}.json")
    rec = {"n": n, "resolution": 0.5, "scene": f"synth.vesicle_scene(box_nm={n / 2.0})", "timings_s": {}}
    t0 = time.time()
    pos, dims = synth.vesicle_scene(box_nm=n / 2.0)
    rec["n_atoms"] = int(len(pos))
    rec["sha_positions"] = sha(pos)
    vox, nbox, _ = vo.voxelate_fma(pos, dims, 0.5)
    assert tuple(np.diagonal(nbox)) == (n, n, n)
    occ = np.zeros((n, n, n), dtype=bool)
    occ[vox[:, 0], vox[:, 1], vox[:, 2]] = True
    del vox
    rec["n_occupied"] = int(occ.sum())
    rec["digest_occupancy"] = digest(occ)
    rec["timings_s"]["bin"] = time.time() - t0

    t0 = time.time()
    closed = vo.morph(occ, "de")
    del occ
    rec["n_closed"] = int(closed.sum())
    rec["digest_closed"] = digest(closed)
    rec["timings_s"]["close"] = time.time() - t0

    t0 = time.time()
    from scipy import ndimage
    st = np.ones((3, 3, 3), dtype=int)
    lab, n_true = ndimage.label(closed, structure=st, output=np.int32)
    neg, n_false = ndimage.label(~closed, structure=st, output=np.int32)
    lab -= neg                                  # == np.where(grid, pos, -neg): the two are disjoint
    del neg
    rec["n_true"], rec["n_false"] = int(n_true), int(n_false)
    rec["digest_labels"] = digest(lab)
    rec["timings_s"]["label"] = time.time() - t0

    t0 = time.time()
    vol = np.zeros(n_true + n_false + 1, dtype=np.int64)
    flat = lab.reshape(-1)
    for s in range(0, flat.size, 1 << 27):
        vol += np.bincount(flat[s:s + (1 << 27)].astype(np.int64) + n_false, minlength=len(vol))
    rec["label_volumes"] = [int(v) for v in vol]          # index = label + n_false
    rec["timings_s"]["volumes"] = time.time() - t0

    t0 = time.time()
    kernels = None
    try:
        import refshim
        kernels = refshim.load_ref_kernels()
    except Exception as e:                                 # noqa: BLE001
        print("reference kernels unavailable:", e)
    if kernels:
        contacts = np.asarray(kernels["find_label_contacts"].find_label_contacts(lab), dtype=np.int32).reshape(-1, 2)
        bridges = np.asarray(kernels["find_bridges"].find_bridges(lab), dtype=np.int32)
        rec["pairs_by"] = "reference Cython kernels (oracle/_ref)"
    else:
        contacts = vo.find_label_contacts_c(lab)
        bridges = vo.find_bridges_c(lab)
        rec["pairs_by"] = "oracle C restatement"
    rec["contacts"] = contacts.tolist()
    rec["bridges"] = bridges.reshape(-1, 5).tolist() if bridges.size else []
    rec["timings_s"]["pairs"] = time.time() - t0
    with open(out_path, "w") as fh:
        json.dump(rec, fh, indent=1)
    print(json.dumps({k: v for k, v in rec.items() if k not in ("contacts", "bridges", "label_volumes")}, indent=1))
    print("contacts", len(rec["contacts"]), "bridges", len(rec["bridges"]), "->", out_path)


if __name__ == "__main__":
    main()
