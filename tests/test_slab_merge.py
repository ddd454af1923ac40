"""Cross-slab merge logic (mdvcontainment_b200/slab.py: merge_slabs) against the oracle on whole grids. CPU only."""
import numpy as np
import pytest

import voxel_oracle as vo
from golden_util import voxel_cases
from mdvcontainment_b200 import host_graph as hg
from mdvcontainment_b200.slab import merge_slabs, merge_slabs_numpy, slab_extents


def plane_pairs_numpy(A, B, sx):
    """NumPy restatement of mdvc_plane_pairs: steps from plane A to the 9 neighbours in plane B, y/z periodic."""
    Y, Z = A.shape
    rows = set()
    iy = np.arange(Y)[:, None]; iz = np.arange(Z)[None, :]
    for dy in (-1, 0, 1):
        for dz in (-1, 0, 1):
            sy = np.where(iy + dy < 0, -1, np.where(iy + dy >= Y, 1, 0))
            sz = np.where(iz + dz < 0, -1, np.where(iz + dz >= Z, 1, 0))
            nb = np.roll(B, (-dy, -dz), axis=(0, 1))
            sy, sz = np.broadcast_to(sy, A.shape), np.broadcast_to(sz, A.shape)
            stack = np.stack([A, nb, np.full(A.shape, sx), sy, sz], -1).reshape(-1, 5)
            rows.update(map(tuple, np.unique(stack, axis=0).tolist()))
    return np.array(sorted(rows), dtype=np.int32).reshape(-1, 5)


def solve_by_slabs(grid, n_slabs):
    X = grid.shape[0]
    ext = slab_extents(X, n_slabs)
    nt, nf, contacts, bridges, volumes, labels = [], [], [], [], [], []
    for x0, x1 in ext:
        lab = vo.label_3d_grid(grid[x0:x1]).astype(np.int32)
        labels.append(lab)
        nt.append(int(max(lab.max(), 0))); nf.append(int(max(-lab.min(), 0)))
        contacts.append(vo.find_label_contacts_c(lab))
        b = vo.find_bridges_c(lab).reshape(-1, 5)
        bridges.append(b[b[:, 2] == 0])                        # y/z faces only: the x faces belong to the boundaries
        vol = np.zeros(nf[-1] + nt[-1] + 1, dtype=np.int64)
        v, c = np.unique(lab, return_counts=True)
        vol[v + nf[-1]] = c
        volumes.append(vol)
    boundary = [plane_pairs_numpy(labels[g][-1], labels[(g + 1) % n_slabs][0], 1 if g == n_slabs - 1 else 0)
                for g in range(n_slabs)]
    m = merge_slabs(nt, nf, boundary, contacts, bridges, volumes)
    ref = merge_slabs_numpy(nt, nf, boundary, contacts, bridges, volumes)     # the NumPy statement of the same rules
    assert (m["n_true"], m["n_false"]) == (ref["n_true"], ref["n_false"])
    assert all(np.array_equal(a, b) for a, b in zip(m["luts"], ref["luts"]))
    assert m["contacts"].shape == ref["contacts"].shape and np.array_equal(m["contacts"], ref["contacts"])
    assert m["bridges"].shape == ref["bridges"].shape and np.array_equal(m["bridges"], ref["bridges"])
    assert np.array_equal(m["volumes"], ref["volumes"])
    glob = np.concatenate([m["luts"][g][labels[g] + nf[g]] for g in range(n_slabs)], axis=0)
    return m, glob


def check(grid, n_slabs):
    m, glob = solve_by_slabs(grid, n_slabs)
    want = vo.label_3d_grid(grid).astype(np.int32)
    assert np.array_equal(glob, want)
    assert (m["n_true"], m["n_false"]) == (max(want.max(), 0), max(-want.min(), 0))
    wc, wb = vo.find_label_contacts_c(want), vo.find_bridges_c(want)
    assert m["contacts"].shape == wc.shape and np.array_equal(m["contacts"], wc)
    assert m["bridges"].shape == wb.shape and np.array_equal(m["bridges"], wb)
    v, c = np.unique(want, return_counts=True)
    assert np.array_equal(m["volumes"][v + m["n_false"]], c) and m["volumes"].sum() == grid.size


def test_extents():
    assert slab_extents(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert slab_extents(8, 8) == [(k, k + 1) for k in range(8)]
    with pytest.raises(ValueError):
        slab_extents(3, 4)


@pytest.mark.parametrize("name", ["complex3D_roll0", "complex3D_roll3", "complex3D_roll11", "complex3D_roll11_closed",
                                  "perlin_32x32x32_s0", "perlin_48x32x16_s2_closed", "rand_33x34x35_p50", "rand_7x70x3_p15",
                                  "salt_40x24x72", "allfalse_4x5x6", "alltrue_4x5x6"])
@pytest.mark.parametrize("n_slabs", [1, 2, 3, 4])
def test_merge_matches_whole_grid_oracle(name, n_slabs):
    g = voxel_cases()[name]["grid"]
    if g.shape[0] < n_slabs:
        pytest.skip("fewer planes than slabs")
    check(g, n_slabs)


def test_merge_random_grids_many_slabs():
    rng = np.random.default_rng(31)
    for _ in range(40):
        shape = (int(rng.integers(2, 14)), int(rng.integers(1, 8)), int(rng.integers(1, 8)))
        g = rng.random(shape) < rng.choice([0.2, 0.5, 0.8])
        check(g, int(rng.integers(1, shape[0] + 1)))
