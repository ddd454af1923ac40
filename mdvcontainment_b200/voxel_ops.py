"""Seam functions of the reference's ``label.py``, ``find_label_contacts.pyx`` and ``find_bridges.pyx``
on the CUDA path: NumPy in / NumPy out with the reference's dtypes, shapes and row order, so the
unchanged host graph code can consume them."""
from __future__ import annotations

import numpy as np
import torch

from . import device as dev


def label_3d_grid(grid):
    """label.py:4-19: 26-connected components of True (+ids) and False (-ids) voxels, no wrap,
    numbered in C-order of their first voxel.  bool [X,Y,Z] -> int32 [X,Y,Z]."""
    grid = np.asarray(grid)
    if grid.dtype != np.bool_:
        raise TypeError("label_3d_grid expects a boolean grid")
    bits = dev.BitGrid.from_bool(grid)
    plan, _ = dev.label_with_retry(bits, want_pairs=False)
    labels = torch.empty(bits.shape, dtype=torch.int32, device="cuda")
    plan.expand(bits, labels, None)
    return dev.to_numpy(labels)


def find_label_contacts(labeled_grid):
    """find_label_contacts.pyx:5-77: sorted unique (min,max) rows, int32 (K,2); (0,2) when empty."""
    L = np.asarray(labeled_grid)
    if L.dtype != np.int32 or L.ndim != 3:
        raise ValueError("Buffer dtype mismatch, expected 'int32_t' 3D array")
    return dev.grid_pairs(L, bridges=False).reshape(-1, 2)


def find_bridges(labeled_grid):
    """find_bridges.pyx:4-89: sorted unique (l1,l2,sx,sy,sz) rows, int32 (K,5); shape (0,) when empty."""
    L = np.asarray(labeled_grid)
    if L.dtype != np.int32 or L.ndim != 3:
        raise ValueError("Buffer dtype mismatch, expected 'int32_t' 3D array")
    return dev.grid_pairs(L, bridges=True)


def create_components_grid(nonp_labeled_grid, component2labels):
    """label.py:21-38: relabel through {component: labels}; labels missing from the dict become 0.
    One LUT gather instead of the reference's O(#labels x N) masked assignments."""
    L = np.asarray(nonp_labeled_grid, dtype=np.int32)
    labs = [int(l) for ls in component2labels.values() for l in ls]
    if not labs:
        return np.zeros_like(L)
    lo, hi = min(labs), max(labs)
    lut = np.zeros(hi - lo + 1, dtype=np.int32)
    for comp, ls in component2labels.items():
        for l in ls:
            lut[int(l) - lo] = comp
    return dev.to_numpy(dev.grid_relabel(L, lut, lo))
