"""On-disk formats: the .gro / .pdb readers and writers of mdvcontainment_b200/universe.py.

The reference reads structures through MDAnalysis (mda_containment.py:324, voxels_to_gro.py:3), which is not in
this image; its one bundled fixture is examples/structures/complex3D.gro (written by MDAnalysis from a 23x10x10
voxel grid with voxels_to_gro).  It holds the grid of the golden case ``complex3D_roll11``, so a universe rebuilt
from that golden grid must equal the universe read from the file."""
import os

import numpy as np
import pytest

from golden_util import voxel_cases
from mdvcontainment_b200 import universe as uni

REF_GRO = "/root/reference/examples/structures/complex3D.gro"


def _same_universe(a, b, pos_tol=0.0):
    assert len(a.atoms) == len(b.atoms)
    if pos_tol:
        assert np.abs(a.atoms.positions - b.atoms.positions).max() <= pos_tol
    else:
        assert np.array_equal(a.atoms.positions, b.atoms.positions)
    assert list(a.atoms.names) == list(b.atoms.names)
    assert np.allclose(a.dimensions, b.dimensions, atol=1e-3)


def test_gro_round_trip(tmp_path):
    grid = voxel_cases()["complex3D_roll0"]["grid"]
    u = uni.voxels_to_universe(grid, 1.0)
    path = tmp_path / "complex.gro"
    uni.write_gro(path, u)
    back = uni.read_gro(path)
    _same_universe(u, back)
    assert back.atoms.positions.dtype == np.float32 and back.dimensions.dtype == np.float32
    sel = back.select_atoms("name True")
    assert len(sel) == int(grid.sum())
    # positions are voxel centres in Angstrom: the occupancy grid comes straight back
    vox = np.floor(sel.positions / 10.0).astype(int)
    occ = np.zeros(grid.shape, dtype=bool)
    occ[tuple(vox.T)] = True
    assert np.array_equal(occ, grid)


def test_voxels_to_gro_matches_the_universe_helper(tmp_path):
    grid = voxel_cases()["complex3D_roll0"]["grid"]
    path = tmp_path / "dump.gro"
    uni.voxels_to_gro(path, grid)
    _same_universe(uni.read_gro(path), uni.voxels_to_universe(grid, 1.0))
    uni.voxels_to_gro(path, grid, scale=0.5, place_in_center=False)
    u = uni.read_gro(path)
    assert np.allclose(u.dimensions[:3], np.array(grid.shape) * 5.0)
    assert np.array_equal(u.atoms.positions[1], np.array([0, 0, 5.0], dtype=np.float32))


def test_triclinic_gro_box(tmp_path):
    dims = np.array([100.0, 90.0, 80.0, 70.0, 80.0, 60.0], dtype=np.float32)
    from mdvcontainment_b200.atomgroup_to_voxels import dim2lattice
    box = dim2lattice(*dims) / 10.0                                   # rows = box vectors, nm
    path = tmp_path / "tric.gro"
    with open(path, "w") as fh:
        fh.write("triclinic\n    1\n    1RES      C    1   1.000   2.000   3.000\n")
        v = box
        fh.write("%10.5f%10.5f%10.5f%10.5f%10.5f%10.5f%10.5f%10.5f%10.5f\n" %
                 (v[0, 0], v[1, 1], v[2, 2], v[0, 1], v[0, 2], v[1, 0], v[1, 2], v[2, 0], v[2, 1]))
    u = uni.load(path)
    assert np.allclose(u.dimensions, dims, atol=2e-3)
    assert np.array_equal(u.atoms.positions, np.array([[10, 20, 30]], dtype=np.float32))


def test_pdb_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    pos = np.round(rng.random((50, 3)) * 40.0, 3).astype(np.float32)
    names = ["C1A", "PO4", "W", "NC3", "GL1"] * 10
    u = uni.Universe(pos, [40, 41, 42, 90, 90, 60], names, ["POPC"] * 50, np.arange(50) // 5 + 1)
    u.add_TopologyAttr("tempfactors", np.arange(50, dtype=float) - 3)
    path = tmp_path / "x.pdb"
    uni.write_pdb(path, u)
    back = uni.load(path)
    _same_universe(u, back)
    assert np.array_equal(back.atoms.tempfactors, u.atoms.tempfactors)
    assert list(back.atoms.resnames) == ["POPC"] * 50 and list(back.atoms.resids) == list(u.atoms.resids)
    with pytest.raises(ValueError):
        uni.load(tmp_path / "x.xyz")


def test_pdb_without_box_is_rejected(tmp_path):
    path = tmp_path / "nobox.pdb"
    path.write_text("ATOM      1  C   RES X   1       1.000   2.000   3.000  1.00  0.00\nEND\n")
    with pytest.raises(ValueError):
        uni.read_pdb(path)


@pytest.mark.reference
def test_bundled_complex3d_fixture_is_read_like_the_golden_grid():
    """examples/structures/complex3D.gro -> same atoms as the universe rebuilt from the golden grid, and our writer
    reproduces the file's atom and box lines byte for byte."""
    assert os.path.exists(REF_GRO)
    u = uni.read_gro(REF_GRO)
    grid = voxel_cases()["complex3D_roll11"]["grid"]                  # the file is gen_data.make_complex_3D(11)
    _same_universe(u, uni.voxels_to_universe(grid, 1.0))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "w.gro")
        uni.voxels_to_gro(out, grid)
        ours = open(out).read().splitlines()
    theirs = open(REF_GRO).read().splitlines()
    assert ours[1:] == theirs[1:]
