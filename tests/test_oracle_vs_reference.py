"""Pins oracle/voxel_oracle.py against the UNMODIFIED reference (build container only)."""
import numpy as np
import pytest

import voxel_oracle as vo
from mdvcontainment_b200.universe import Universe

pytestmark = pytest.mark.reference


def random_grids(n, seed=0, maxdim=9):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        shape = tuple(int(v) for v in rng.integers(1, maxdim, 3))
        yield rng.random(shape) < rng.choice([0.1, 0.3, 0.5, 0.7, 0.9])


def test_morph_matches_reference(ref):
    from mdvcontainment.atomgroup_to_voxels import morph_voxels
    for g in random_grids(120, 1):
        for ms in ("d", "e", "de", "ed", "dde", "deed"):
            want = morph_voxels(g.copy(), ms)
            assert np.array_equal(vo.morph(g, ms), want)
            assert np.array_equal(vo.morph_c(g, ms), want)
    with pytest.raises(ValueError):
        vo.morph(np.zeros((2, 2, 2), bool), "dx")


def test_label_contacts_bridges_match_reference(ref):
    from mdvcontainment.label import label_3d_grid
    from mdvcontainment.find_label_contacts import find_label_contacts
    from mdvcontainment.find_bridges import find_bridges
    for g in random_grids(150, 2, maxdim=12):
        want = label_3d_grid(g).astype(np.int32)
        got = vo.label_3d_grid(g)
        assert np.array_equal(got, want)
        lab_c, nt, nf = vo.label_3d_grid_c(g)
        assert np.array_equal(lab_c, want)
        assert nt == max(want.max(), 0) and nf == max(-want.min(), 0)
        wc = find_label_contacts(want)
        for c in (vo.find_label_contacts(want), vo.find_label_contacts_c(want)):
            assert c.shape == wc.shape and c.dtype == np.int32 and np.array_equal(c, wc)
        wb = find_bridges(want)
        for b in (vo.find_bridges(want), vo.find_bridges_c(want)):
            assert b.shape == wb.shape and b.dtype == np.int32 and np.array_equal(b, wb)


def test_components_grid_and_counts(ref):
    from mdvcontainment.label import create_components_grid, label_3d_grid
    rng = np.random.default_rng(3)
    for g in random_grids(40, 4, maxdim=10):
        lab = label_3d_grid(g).astype(np.int32)
        labs = np.unique(lab)
        comps = rng.integers(-3, 4, len(labs))
        c2l = {}
        for l, c in zip(labs, comps):
            if c != 0:
                c2l.setdefault(int(c), []).append(int(l))
        c2l = {k: tuple(v) for k, v in c2l.items()}
        assert np.array_equal(vo.create_components_grid(lab, c2l), create_components_grid(lab, c2l))


@pytest.mark.parametrize("dims", [
    (100.0, 80.0, 60.0, 90.0, 90.0, 90.0),
    (97.3, 97.3, 45.1, 90.0, 90.0, 90.0),
    (120.0, 120.0, 90.0, 90.0, 90.0, 60.0),
    (100.0, 100.0, 100.0, 60.0, 60.0, 90.0),
    (64.5, 71.25, 80.0, 75.0, 80.0, 85.0),
])
@pytest.mark.parametrize("res", [0.5, 1.0, 0.37])
def test_binning_matches_reference(ref, dims, res):
    from mdvcontainment.atomgroup_to_voxels import create_voxels
    rng = np.random.default_rng(5)
    n = 20000
    dims32 = np.array(dims, dtype=np.float32)
    pos = (rng.random((n, 3)) * 3 - 1) * dims32[:3] * 1.0          # inside and outside the box
    pos[: n // 2] = np.round(pos[: n // 2] / (10 * res)) * (10 * res)   # exactly on voxel faces
    pos = np.round(pos, 2).astype(np.float32)
    u = Universe(pos.copy(), dims32)
    # two successive passes, as Containment does (mda_containment.py:385-388): the write-back matters
    p = pos.copy()
    for _ in range(3):
        grid, mapping = create_voxels(u.atoms, res, return_mapping=True)
        vox_np, nbox, new_np = vo.voxelate_numpy(p, dims32, res)
        vox_fm, _, new_fm = vo.voxelate_fma(p, dims32, res)
        assert np.array_equal(mapping["atom_voxels"], vox_np.astype(np.int32))
        assert np.array_equal(vox_np, vox_fm)
        assert np.array_equal(u.atoms.positions.view(np.uint32), new_np.view(np.uint32))
        assert np.array_equal(new_np.view(np.uint32), new_fm.view(np.uint32))
        assert np.array_equal(vo.occupancy(vox_fm, nbox), grid)
        assert (vox_fm >= 0).all() and (vox_fm < np.diagonal(nbox)).all()
        p = new_fm


def test_mapping_matches_reference_when_hash_is_collision_free(ref):
    from mdvcontainment.atomgroup_to_voxels import _create_efficient_mapping
    from mdvcontainment.atoms_voxels_mapping import voxels2atomgroup_cy
    rng = np.random.default_rng(6)
    tested = 0
    for _ in range(20):
        shape = rng.integers(2, 24, 3)
        n = int(rng.integers(1, 3000))
        vox = np.stack([rng.integers(0, s, n) for s in shape], 1)
        ix = rng.permutation(10 * n)[:n].astype(np.int64)
        if vo.reference_hash_collides(vox):
            continue
        tested += 1
        want = _create_efficient_mapping(vox, ix)
        got = vo.create_mapping(vox, ix)
        assert set(want["voxel_to_atoms"]) == set(got["voxel_to_atoms"])
        for k, v in want["voxel_to_atoms"].items():
            assert np.array_equal(v, got["voxel_to_atoms"][k])
        q = np.stack([rng.integers(0, s, 500) for s in shape], 1).astype(np.int32)
        assert np.array_equal(voxels2atomgroup_cy(q, want["voxel_to_atoms"], want["atom_indices"]),
                              vo.voxels2atoms(q, got))
    assert tested >= 10


def test_reference_hash_restatement_detects_the_known_collisions(ref):
    # 4 colliding voxels in a full 64^3 grid
    from mdvcontainment.atomgroup_to_voxels import _create_efficient_mapping
    idx = np.indices((64, 64, 64)).reshape(3, -1).T
    m = _create_efficient_mapping(idx, np.arange(len(idx)))
    lost = len(idx) - len(m["voxel_to_atoms"])
    h = vo.reference_voxel_hash(idx)
    assert lost == len(idx) - len(np.unique(h)) and lost > 0
