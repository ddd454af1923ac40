"""End-to-end parity of Containment / VoxelContainment (CUDA path) with the reference's recorded results."""
import contextlib
import io
import json

import numpy as np
import pytest

import e2e_inputs
from golden_util import e2e_cases, sha, voxel_cases

pytestmark = pytest.mark.gpu

E2E = e2e_cases()
VOX = voxel_cases()


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = fn(*a, **k)
    return out, buf.getvalue()


@pytest.mark.parametrize("name", sorted(VOX))
def test_voxel_containment_matches_reference(name):
    from mdvcontainment_b200 import VoxelContainment
    c = VOX[name]
    vc = VoxelContainment(c["grid"])
    assert str(vc) == c["str"]
    assert [int(n) for n in vc.containment_graph.nodes] == c["nodes"]
    assert [[int(a), int(b)] for a, b in vc.containment_graph.edges()] == c["edges"]
    assert {str(int(k)): int(v) for k, v in vc.component_ranks.items()} == c["ranks"]
    assert {str(int(k)): int(v) for k, v in vc.voxel_counts.items()} == c["counts"]
    assert all(isinstance(k, np.int32) for k in vc.voxel_counts) and all(isinstance(v, np.int64) for v in vc.voxel_counts.values())
    comps = vc.components_grid
    assert comps.dtype == np.int32 and sha(comps) == c["sha_components"]
    assert sha(vc.nonp_labeled_grid_device.cpu().numpy()) == c["sha_labels"]
    assert vc.grid is c["grid"]
    # array queries
    nodes = vc.nodes[:3]
    mask = vc.get_voxel_mask(nodes)
    assert np.array_equal(mask, np.isin(comps, nodes))
    assert np.array_equal(vc.get_voxel_positions(nodes), np.array(np.where(mask)).T)
    assert vc.get_counts(comps) == dict(zip(*np.unique(comps, return_counts=True)))


@pytest.mark.parametrize("key", sorted(E2E))
def test_containment_matches_reference(key):
    from mdvcontainment_b200 import Containment
    rec = E2E[key]
    name = key.split("|")[0]
    kw = rec["kwargs"]
    u, sel = e2e_inputs.build(name, VOX["complex3D_roll11"]["grid"] if name == "complex3D" else None)
    assert sha(u.atoms.positions) == rec["sha_input_positions"]
    c, printed = quiet(Containment, sel, **kw)
    vc = c.voxel_containment
    assert printed == rec["printed"]
    assert str(c) == rec["str"]
    assert str(vc) == rec["str_voxels"]
    assert list(c.boolean_grid.shape) == rec["grid_shape"] and sha(c.boolean_grid, np.uint8) == rec["sha_grid"]
    assert sha(vc.components_grid) == rec["sha_components"]
    assert sha(u.atoms.positions) == rec["sha_positions_after"]
    assert float(c.voxel_volume) == rec["voxel_volume"]
    assert [int(n) for n in vc.nodes] == rec["nodes"]
    assert [[int(a), int(b)] for a, b in vc.containment_graph.edges()] == rec["edges"]
    assert {str(int(k)): int(v) for k, v in vc.component_ranks.items()} == rec["ranks"]
    assert {str(int(k)): int(v) for k, v in vc.voxel_counts.items()} == rec["counts"]
    if kw.get("no_mapping", False):
        with pytest.raises(ValueError):
            c.get_atomgroup_from_nodes([1])
        with pytest.raises(ValueError):
            c.set_betafactors()
        return
    tf = u.atoms.tempfactors
    assert sha(tf, np.float64) == rec["sha_tempfactors"]
    for node, (na, sa, nb, sb) in rec["atoms_per_node"].items():
        a = c.get_atomgroup_from_nodes([int(node)])
        b = c.get_atomgroup_from_nodes([int(node)], containment=True)
        assert (len(a), sha(a.ix), len(b), sha(b.ix)) == (na, sa, nb, sb), node
        # the generic voxel-list route gives the same atoms
        q = vc.get_voxel_positions([int(node)])
        assert np.array_equal(c.get_atomgroup_from_voxel_positions(q).ix, a.ix)
    if "view_str" in rec:
        view = c.node_view(min_size=rec["view_min_size"])
        assert str(view) == rec["view_str"]
        assert [int(n) for n in view.nodes] == rec["view_nodes"]
        for n, (cnt, h) in rec["view_atoms"].items():
            ag = view.get_atomgroup_from_nodes([int(n)])
            assert (len(ag), sha(ag.ix)) == (cnt, h)
        _, msg = quiet(view.set_betafactors)
        assert "VIEW" in msg
        vals = set(np.unique(u.atoms.tempfactors).astype(int).tolist())
        assert vals <= set(rec["view_nodes"]) | {0}


def test_constructor_validation_matches_reference():
    from mdvcontainment_b200 import Containment
    u, sel = e2e_inputs.build("complex3D", VOX["complex3D_roll11"]["grid"])
    with pytest.raises(AssertionError, match="Atomgroup must be"):
        Containment("not an atomgroup", 1)
    with pytest.raises(AssertionError, match="Resolution must be"):
        Containment(sel, "1")
    with pytest.raises(AssertionError, match="Morph strings"):
        Containment(sel, 1, morph="dx")
    with pytest.raises(ValueError, match="betafactors='True' requires"):
        Containment(sel, 1, no_mapping=True)


def test_seam_functions_are_numpy_in_numpy_out():
    from mdvcontainment_b200.voxel_ops import create_components_grid, find_bridges, find_label_contacts, label_3d_grid
    from mdvcontainment_b200.atomgroup_to_voxels import morph_voxels, create_voxels
    c = VOX["complex3D_roll3"]
    lab = label_3d_grid(c["grid"])
    assert lab.dtype == np.int32 and np.array_equal(lab, c["labels"])
    assert np.array_equal(find_label_contacts(lab), c["contacts"])
    assert np.array_equal(find_bridges(lab), c["bridges"])
    assert np.array_equal(morph_voxels(VOX["complex3D_roll11"]["grid"], "de"), VOX["complex3D_roll11_closed"]["grid"])
    with pytest.raises(ValueError, match="Unknown morph operation 'x'"):
        morph_voxels(c["grid"], "dx")
    with pytest.raises(ValueError):
        find_bridges(lab.astype(np.int64))
    u, sel = e2e_inputs.build("complex3D", VOX["complex3D_roll11"]["grid"])
    grid, mapping = create_voxels(sel, 1, return_mapping=True)
    assert grid.dtype == np.bool_ and np.array_equal(grid, VOX["complex3D_roll11"]["grid"])
    assert set(mapping.keys()) == {"atom_voxels", "voxel_to_atoms", "atom_indices"}
    assert sum(len(v) for v in mapping["voxel_to_atoms"].values()) == len(sel)


def test_frame_pipeline_and_trajectory_driver_match_containment():
    """The reusable per-shape pipeline (bench / trajectory path) gives the same DAG as Containment."""
    import torch
    from mdvcontainment_b200 import Containment, synth
    from mdvcontainment_b200.pipeline import FramePipeline
    from mdvcontainment_b200.trajectory import run_trajectory
    from mdvcontainment_b200.universe import Universe
    frames = [synth.trajectory_frame(f, box_nm=24.0) for f in range(3)]
    dims = frames[0][1]
    pipe = FramePipeline(dims, 0.5, closing=True)
    res = run_trajectory([p for p, _ in frames], dims, 0.5, closing=True)
    for f, (pos, _) in enumerate(frames):
        u = Universe(pos.copy(), dims)
        c, _ = quiet(Containment, u.atoms, 0.5, closing=True, no_mapping=True, betafactors=False)
        vc = c.voxel_containment
        out = pipe.run(torch.from_numpy(pos).cuda())
        assert str(out) == str(vc)
        assert np.array_equal(out.components_dev.cpu().numpy(), vc.components_grid)
        assert np.array_equal(out.labels_dev.cpu().numpy(), vc.nonp_labeled_grid_device.cpu().numpy())
        assert res[f]["str"] == str(vc)
        assert res[f]["counts"] == {int(k): int(v) for k, v in vc.voxel_counts.items()}


def test_translation_invariance_of_the_containment_hierarchy():
    """README.md:12 of the reference: rolling the periodic grid must not change the hierarchy (up to renumbering)."""
    import networkx as nx
    from mdvcontainment_b200 import VoxelContainment
    rng = np.random.default_rng(5)
    base = VOX["perlin_48x32x16_s2_closed"]["grid"]

    def signature(vc):
        g = vc.containment_graph
        lab = {n: (int(vc.voxel_counts[n]), int(vc.component_ranks[n]), int(n) > 0) for n in g.nodes}
        return lab, g

    ref_lab, ref_g = signature(VoxelContainment(base))
    for _ in range(4):
        shift = tuple(int(rng.integers(0, s)) for s in base.shape)
        vc = VoxelContainment(np.roll(base, shift, axis=(0, 1, 2)))
        lab, g = signature(vc)
        assert sorted(lab.values()) == sorted(ref_lab.values())
        nx.set_node_attributes(g, lab, "sig"); nx.set_node_attributes(ref_g, ref_lab, "sig")
        assert nx.is_isomorphic(nx.DiGraph(g), nx.DiGraph(ref_g), node_match=lambda a, b: a["sig"] == b["sig"])


def test_pipeline_memoised_graph_stage_is_keyed_by_topology():
    """The pipeline memoises the host graph stage on (label counts, contacts, bridges): alternating scenes with
    different topologies through ONE pipeline must give what a fresh pipeline gives for each."""
    import torch
    from mdvcontainment_b200 import synth
    from mdvcontainment_b200.pipeline import FramePipeline
    a, dims = synth.vesicle_scene(box_nm=24.0, seed=3)
    rng = np.random.default_rng(4)
    b = (rng.random((4000, 3)) * dims[:3]).astype(np.float32)            # sparse random beads: many small components
    c = a[: len(a) // 3]                                                 # only part of the slab: a different hierarchy
    pipe = FramePipeline(dims, 0.5, closing=True)
    seen = {}
    for name, pos in [("a", a), ("b", b), ("a", a), ("c", c), ("b", b), ("c", c), ("a", a)]:
        out = pipe.run(torch.from_numpy(pos).cuda())
        got = (str(out), {int(k): int(v) for k, v in out.voxel_counts.items()},
               {int(k): int(v) for k, v in out.component_ranks.items()}, sha(out.components_dev.cpu().numpy()))
        if name not in seen:
            fresh = FramePipeline(dims, 0.5, closing=True).run(torch.from_numpy(pos).cuda())
            seen[name] = (str(fresh), {int(k): int(v) for k, v in fresh.voxel_counts.items()},
                          {int(k): int(v) for k, v in fresh.component_ranks.items()}, sha(fresh.components_dev.cpu().numpy()))
        assert got == seen[name], name
    assert len({v[0] for v in seen.values()}) == 3                       # the three scenes really differ
    assert 1 <= len(pipe._solved) <= 3
