"""Oracle vs the committed golden vectors and SURVEY.md Appendix D (runs on any box, CPU only)."""
import numpy as np
import pytest

import voxel_oracle as vo
from golden_util import APPENDIX_D, binning_cases, sha, voxel_cases

CASES = voxel_cases()


def test_goldens_agree_with_appendix_d():
    for name, (g, l, c, b, k) in APPENDIX_D.items():
        m = CASES[name]
        assert (m["sha_grid"], m["sha_labels"], m["sha_contacts"], m["sha_bridges"], m["sha_components"]) == (g, l, c, b, k), name


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_voxel_path_matches_golden(name):
    c = CASES[name]
    g = c["grid"]
    assert sha(g, np.uint8) == c["sha_grid"]
    labels = vo.label_3d_grid(g).astype(np.int32)
    assert sha(labels) == c["sha_labels"]
    lab_c, nt, nf = vo.label_3d_grid_c(g)
    assert np.array_equal(lab_c, labels) and (nt, nf) == (c["n_pos"], c["n_neg"])
    if c["has_arrays"]:
        assert np.array_equal(labels, c["labels"])
    for fn in (vo.find_label_contacts_c, vo.find_label_contacts) if g.size <= 40000 else (vo.find_label_contacts_c,):
        got = fn(labels)
        assert got.shape == c["contacts"].shape and np.array_equal(got, c["contacts"])
    for fn in (vo.find_bridges_c, vo.find_bridges) if g.size <= 40000 else (vo.find_bridges_c,):
        got = fn(labels)
        assert got.shape == c["bridges"].shape and np.array_equal(got, c["bridges"])


def test_closed_cases_are_the_closing_of_their_open_case():
    for name in CASES:
        if name.endswith("_closed") and name[:-7] in CASES:
            assert np.array_equal(vo.morph(CASES[name[:-7]]["grid"], "de"), CASES[name]["grid"])
            assert np.array_equal(vo.morph_c(CASES[name[:-7]]["grid"], "de"), CASES[name]["grid"])


@pytest.mark.parametrize("name", sorted(binning_cases()))
def test_oracle_binning_matches_golden(name):
    c = binning_cases()[name]
    pos = c["positions"]
    for p in (1, 2, 3):
        vox, nbox, newpos = vo.voxelate_fma(pos, c["dimensions"], c["resolution"])
        assert np.array_equal(vox.astype(np.int32), c[f"pass{p}_voxels"])
        assert np.array_equal(newpos.view(np.uint32), c[f"pass{p}_positions"].view(np.uint32))
        assert np.array_equal(vo.occupancy(vox, nbox), c[f"pass{p}_grid"])
        pos = newpos


def test_numpy_matmul_equals_fma_chain_on_this_host():
    """Informational pin: the host BLAS evaluates the [n,3]@[3,3] float64
    product as the FMA chain the CUDA kernel uses.  If a different host BLAS breaks this, the
    goldens (generated where it held) remain the contract."""
    for name, c in binning_cases().items():
        a = vo.voxelate_numpy(c["positions"], c["dimensions"], c["resolution"])
        b = vo.voxelate_fma(c["positions"], c["dimensions"], c["resolution"])
        if not (np.array_equal(a[0], b[0]) and np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))):
            pytest.xfail(f"host BLAS does not evaluate matmul as an FMA chain ({name})")
