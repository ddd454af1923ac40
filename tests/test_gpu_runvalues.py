"""Per-atom / per-voxel values straight from the run tables (mdvc_frame_atom_values, mdvc_frame_values_at) against
look-ups in the dense grids, and the lazy materialisation of the components grid in the drop-in API."""
import numpy as np
import pytest

import e2e_inputs
from golden_util import voxel_cases

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
VOX = voxel_cases()


@pytest.mark.parametrize("name", ["complex3D_roll3", "perlin_48x32x16_s2_closed", "rand_33x34x35_p50", "rand_12x11x65_p15",
                                  "salt48", "rand_1x1x5_p50", "allfalse_4x5x6", "rand_2x3x70_p85"])
def test_run_table_values_equal_dense_grid_lookups(name):
    from mdvcontainment_b200 import VoxelContainment
    g = VOX[name]["grid"]
    vc = VoxelContainment(g)
    fr = vc._frame
    assert fr.components_dev is None                       # nothing dense was written by the constructor
    rng = np.random.default_rng(3)
    n = 5000
    vox = np.stack([rng.integers(-1, s + 1, n) for s in g.shape], 1).astype(np.int32)     # some outside the grid
    inside = ((vox >= 0) & (vox < np.array(g.shape))).all(1)
    vox_d = torch.from_numpy(vox).cuda()
    for which, dense in ((0, vc.nonp_labeled_grid_device), (1, vc.components_grid_device)):
        dense = dense.cpu().numpy()
        want = np.zeros(n)
        want[inside] = dense[tuple(vox[inside].T)]
        got = fr.plan.atom_values(vox_d, which).cpu().numpy()
        assert got.dtype == np.float64 and np.array_equal(got, want), (name, which)
        keys = np.sort(rng.integers(0, g.size, 3000)).astype(np.int64)
        got_k = fr.plan.values_at(torch.from_numpy(keys).cuda(), which).cpu().numpy()
        assert got_k.dtype == np.int32 and np.array_equal(got_k, dense.reshape(-1)[keys]), (name, which)
    assert fr.plan.atom_values(vox_d[:0], 1).shape == (0,)


def test_constructor_defers_the_dense_components_grid():
    """Containment with mapping and tempfactors answers set_betafactors / get_atomgroup_from_nodes from the run tables;
    the 4 B/voxel grid appears only when components_grid (or a mask / position query) is read."""
    import contextlib
    import io
    from mdvcontainment_b200 import Containment
    u, sel = e2e_inputs.build("vesicle32")
    with contextlib.redirect_stdout(io.StringIO()):
        c = Containment(sel, 0.5, closing=True)
    fr = c.voxel_containment._frame
    assert fr.components_dev is None and fr.labels_dev is None
    nodes = c.voxel_containment.nodes
    atoms = c.get_atomgroup_from_nodes(nodes[:1], containment=True)
    assert fr.components_dev is None
    tf = u.atoms.tempfactors.copy()
    grid = c.voxel_containment.components_grid                  # now it is written
    assert fr.components_dev is not None and grid.dtype == np.int32
    # tempfactors == components[voxel of atom]; node atoms == atoms whose voxel is in the node set (voxel order)
    vox = c.voxel2atom["atom_voxels"]
    assert np.array_equal(tf, grid[tuple(vox.T)].astype(np.float64))
    down = c.voxel_containment.get_downstream_nodes(nodes[:1])
    lin = (vox[:, 0].astype(np.int64) * grid.shape[1] + vox[:, 1]) * grid.shape[2] + vox[:, 2]
    hit = np.isin(grid[tuple(vox.T)], down)
    order = np.argsort(lin[hit], kind="stable")
    assert np.array_equal(atoms.ix, u.atoms.ix[np.flatnonzero(hit)[order]])
