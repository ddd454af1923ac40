"""Builders of the end-to-end inputs shared by tests/golden/make_golden.py and the GPU tests."""
import numpy as np

from mdvcontainment_b200 import synth
from mdvcontainment_b200.universe import Universe, voxels_to_universe


def complex3d(grid):
    u = voxels_to_universe(grid, 1.0)
    return u, u.select_atoms("name True")


def vesicle(box):
    pos, dims = synth.vesicle_scene(box_nm=box, seed=7)
    rng = np.random.default_rng(99)
    water = np.round(rng.random((len(pos) // 2, 3)) * dims[:3], 2).astype(np.float32)
    names = ["C"] * len(pos) + ["W"] * len(water)
    u = Universe(np.concatenate([pos, water]), dims, names)
    return u, u.select_atoms("name C")


def bicelle():
    pos, dims, names = synth.bicelle_scene()
    u = Universe(pos, dims, names)
    return u, u.select_atoms("name C")


def npt_dimensions(f, box_nm=32.0):
    """Box of frame f of the NPT trajectory set: isotropic breathing of +-1.2 % (A, float32).  Shared by
    tests/golden/make_config_golden.py and the GPU test, so both sides bin with the same float32 box."""
    scale = 1.0 + 0.012 * np.sin(2 * np.pi * f / 7.0)
    return np.array([box_nm * 10 * scale] * 3 + [90, 90, 90], dtype=np.float32)


RUNS = {
    "complex3D": [dict(resolution=1, closing=False), dict(resolution=1, closing=True),
                  dict(resolution=0.5, closing=True), dict(resolution=1, morph="dde"),
                  dict(resolution=1, closing=False, no_mapping=True, betafactors=False),
                  dict(resolution=1, slab=True)],
    "vesicle32": [dict(resolution=0.5, closing=True), dict(resolution=0.5, closing=False),
                  dict(resolution=1.0, closing=False)],
    "vesicle64": [dict(resolution=0.5, closing=True)],
    "bicelle": [dict(resolution=0.5, closing=True), dict(resolution=1.0)],
}


def build(name, complex_grid=None):
    if name == "complex3D":
        return complex3d(complex_grid)
    if name == "vesicle32":
        return vesicle(32.0)
    if name == "vesicle64":
        return vesicle(64.0)
    if name == "bicelle":
        return bicelle()
    raise KeyError(name)
