"""BASELINE.json's configs 2, 4 and 5 on the CUDA path against answers recorded from the UNMODIFIED reference
(tests/golden/config_cases.json, written by tests/golden/make_config_golden.py) and, where the reference cannot
run at all (config 5: more voxels / components than it handles), against the single-GPU path.
"""
import json
import os

import numpy as np
import pytest

from e2e_inputs import npt_dimensions
from golden_util import GOLD, sha, tensor_digest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

with open(os.path.join(GOLD, "config_cases.json")) as fh:
    CFG = json.load(fh)


# ------------------------------------------------------------------------------------------------
# config 2: hollow cylinder at 512^3, closing=True
# ------------------------------------------------------------------------------------------------
def test_config2_hollow_cylinder_512_known_answer():
    """SURVEY.md Appendix D / 8(d) config 2 (generator anchor gen_data.py:13-14): the reference's answer is
    [-2: 96270336: 3] -> [1: 21088256: 1] -> [-1: 16859136: 1]; every stage array against the reference's own
    (digests of the closed grid, the label grid and the components grid; contacts and bridges verbatim)."""
    from mdvcontainment_b200 import VoxelContainment, device as dev, synth
    rec = CFG["cylinder_512"]
    n = 512
    ring2d = synth.hollow_cylinder_grid(n, 1)[:, :, 0]
    bits = dev.BitGrid.from_bool(torch.from_numpy(ring2d).cuda()[:, :, None].expand(n, n, n).contiguous())
    assert bits.count() == rec["n_true_input"]
    closed = dev.morph(bits, "de")
    assert closed.count() == rec["n_true_closed"]
    assert tensor_digest(closed.to_bool()) == rec["digest_closed"]
    vc = VoxelContainment(closed)
    text = str(vc)
    assert text == rec["str"]
    lines = text.splitlines()
    assert lines[0] == "Containment Graph with 3 components (component: nvoxels: rank):"
    assert [ln.strip("└─ │") for ln in lines[1:]] == ["[-2: 96270336: 3]", "[1: 21088256: 1]", "[-1: 16859136: 1]"]
    assert {str(int(k)): int(v) for k, v in vc.voxel_counts.items()} == rec["counts"]
    assert {str(int(k)): int(v) for k, v in vc.component_ranks.items()} == rec["ranks"]
    assert [int(x) for x in vc.nodes] == rec["nodes"]
    assert [[int(a), int(b)] for a, b in vc.containment_graph.edges()] == rec["edges"]
    fr = vc._frame
    assert np.asarray(fr.contacts).reshape(-1, 2).tolist() == rec["contacts"]
    assert np.asarray(fr.bridges).reshape(-1, 5).tolist() == rec["bridges"]
    assert tensor_digest(vc.nonp_labeled_grid_device) == rec["digest_labels"]
    assert tensor_digest(vc.components_grid_device) == rec["digest_components"]


# ------------------------------------------------------------------------------------------------
# config 4: trajectory driver against the reference, frame by frame
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["nvt", "npt"])
def test_run_trajectory_matches_the_reference_frame_by_frame(kind):
    """run_trajectory (pinned staging, pipelines in flight, memoised graph stage) against the reference's
    Containment(sel, 0.5, closing=True) on the same frames.  The NPT set changes the box every frame -- the transform
    always, the grid shape for some frames -- which the reference handles by reading atomgroup.dimensions in every
    constructor (atomgroup_to_voxels.py:112-131)."""
    from mdvcontainment_b200 import synth
    from mdvcontainment_b200.pipeline import FramePipeline
    from mdvcontainment_b200.trajectory import run_trajectory
    rec = CFG["trajectory_" + kind]
    frames = sorted(int(f) for f in rec["frames"])
    box = rec["box_nm"]
    positions = [synth.trajectory_frame(f, box_nm=box)[0] for f in frames]
    fixed = synth.trajectory_frame(0, box_nm=box)[1]
    dims = fixed if kind == "nvt" else np.stack([npt_dimensions(f, box) for f in frames])
    res = run_trajectory(positions, dims, 0.5, closing=True)
    shapes = set()
    for k, f in enumerate(frames):
        want = rec["frames"][str(f)]
        assert res[k]["str"] == want["str"], f
        assert res[k]["nodes"] == want["nodes"] and [list(e) for e in res[k]["edges"]] == want["edges"], f
        assert {str(a): b for a, b in res[k]["ranks"].items()} == want["ranks"], f
        assert {str(a): b for a, b in res[k]["counts"].items()} == want["counts"], f
        shapes.add(tuple(want["grid_shape"]))
        # the dense grids of the same frame through a fresh pipeline
        d = fixed if kind == "nvt" else dims[k]
        pipe = FramePipeline(d, 0.5, closing=True)
        out = pipe.run(torch.from_numpy(positions[k]).cuda())
        assert pipe.shape == tuple(want["grid_shape"])
        assert sha(out.bits.to_bool().cpu().numpy(), np.uint8) == want["sha_grid"], f
        assert sha(out.components_dev.cpu().numpy()) == want["sha_components"], f
    if kind == "npt":
        assert len(shapes) > 1                            # the box change really moved the grid shape


def test_pipeline_set_dimensions_keeps_buffers_and_changes_the_transform():
    from mdvcontainment_b200 import synth
    from mdvcontainment_b200.pipeline import FramePipeline
    rec = CFG["trajectory_npt"]
    same = [int(f) for f, r in rec["frames"].items() if r["grid_shape"] == rec["frames"]["0"]["grid_shape"]]
    other = [int(f) for f, r in rec["frames"].items() if r["grid_shape"] != rec["frames"]["0"]["grid_shape"]]
    assert len(same) > 1 and other
    pipe = FramePipeline(npt_dimensions(same[0]), 0.5, closing=True)
    ws = pipe.plan.ws.data_ptr()
    for f in sorted(same) * 2:                            # second round replays / recaptures the CUDA graph
        assert pipe.set_dimensions(npt_dimensions(f))
        out = pipe.run(torch.from_numpy(synth.trajectory_frame(f, box_nm=32.0)[0]).cuda())
        assert str(out) == rec["frames"][str(f)]["str"], f
        assert sha(out.components_dev.cpu().numpy()) == rec["frames"][str(f)]["sha_components"], f
    assert pipe.plan.ws.data_ptr() == ws
    assert not pipe.set_dimensions(npt_dimensions(other[0]))        # another grid shape needs another pipeline


# ------------------------------------------------------------------------------------------------
# many labels: host graph stage inside the device path
# ------------------------------------------------------------------------------------------------
def test_many_label_grid_128_matches_the_reference_graph_stage():
    """BASELINE.md section 2's pathological case (128^3 Bernoulli(0.05), 48 k labels): device labels / contacts /
    bridges + the array host graph stage against the reference's recorded numbering, ranks and DAG."""
    from mdvcontainment_b200 import VoxelContainment
    rec = CFG["many_labels_128"]
    rng = np.random.default_rng(rec["seed"])
    grid = rng.random((rec["n"],) * 3) < rec["p"]
    assert sha(grid, np.uint8) == rec["sha_grid"]
    vc = VoxelContainment(grid)
    fr = vc._frame
    assert len(fr.unique_labels) == rec["n_labels"]
    assert sha(fr.contacts) == rec["sha_contacts"] and sha(fr.bridges) == rec["sha_bridges"]
    assert sha(fr.lut) == rec["sha_lut"]
    assert sha(np.array([int(r) for r in vc.component_ranks.values()], dtype=np.int64)) == rec["sha_ranks"]
    assert sha(np.array([int(c) for c in vc.component_ranks], dtype=np.int64)) == rec["sha_component_order"]
    g = vc.containment_graph
    assert sha(np.array([int(x) for x in g.nodes], dtype=np.int64)) == rec["dag"]["sha_nodes"]
    assert sha(np.array([[int(a), int(b)] for a, b in g.edges()], dtype=np.int64).reshape(-1, 2)) == rec["dag"]["sha_edges"]
    ccg = vc.component_contact_graph
    assert sha(np.array([int(x) for x in ccg.nodes], dtype=np.int64)) == rec["sha_ccg_nodes"]
    assert sha(np.array([int(v) for x in ccg.nodes for v in ccg.neighbors(x)], dtype=np.int64)) == rec["sha_ccg_adjacency"]
    assert sum(int(v) for v in vc.voxel_counts.values()) == grid.size
    assert vc.timings["host_graph_s"] < 2.0               # the reference: 22.8 s


# ------------------------------------------------------------------------------------------------
# config 5: slab decomposition at 1024^3 == single GPU (many labels)
# ------------------------------------------------------------------------------------------------
def test_slab_decomposition_equals_single_gpu_at_1024_with_many_labels():
    """one 1024^3 many-label grid (periodic gyroid XOR Bernoulli(1e-3) noise, about
    a million labels -- more components than the reference's orientation loop accepts) analysed whole on one GPU and
    as 4 x-slabs (LocalGather: the ranks emulated on this GPU, same code path as NCCL apart from the gather)."""
    from mdvcontainment_b200 import device as dev, synth
    from mdvcontainment_b200.pipeline import FramePipeline
    from mdvcontainment_b200.slab import LocalGather, analyse_slabs, slab_extents
    if torch.cuda.mem_get_info()[0] < 60e9:
        pytest.skip("needs about 60 GB of free device memory")
    n = 1024
    shape = (n, n, n)
    whole = synth.gyroid_salt_planes(0, n, shape, salt=1e-3)
    dims = np.array([n * 5.0] * 3 + [90, 90, 90], dtype=np.float32)          # 0.5 nm voxels
    pipe = FramePipeline(dims, 0.5, closing=False)
    assert pipe.shape == shape
    single = pipe.analyse(whole)
    torch.cuda.synchronize()
    assert single.n_true + single.n_false > 500_000
    d_lab, d_comp = tensor_digest(single.labels_dev), tensor_digest(single.components_dev)
    counts = dict(zip(single.count_ids.tolist(), single.count_values.tolist()))
    assert sum(counts.values()) == n ** 3
    s_nodes, s_edges = single.dag.nodes.copy(), (single.dag.edge_src.copy(), single.dag.edge_dst.copy())
    s_contacts, s_bridges = single.contacts.copy(), single.bridges.copy()
    del single, pipe
    torch.cuda.empty_cache()

    ext = slab_extents(n, 4)
    bits = {g: dev.BitGrid((x1 - x0, n, n), whole.words[x0:x1]) for g, (x0, x1) in enumerate(ext)}
    res = analyse_slabs(bits, shape, LocalGather(4))
    torch.cuda.synchronize()
    assert np.array_equal(res.contacts, s_contacts) and np.array_equal(res.bridges, s_bridges)
    assert tensor_digest(torch.cat([res.labels[g] for g in range(4)])) == d_lab
    assert tensor_digest(torch.cat([res.components[g] for g in range(4)])) == d_comp
    assert dict(zip(res.count_ids.tolist(), res.count_values.tolist())) == counts
    assert np.array_equal(res.dag.nodes, s_nodes)
    assert np.array_equal(res.dag.edge_src, s_edges[0]) and np.array_equal(res.dag.edge_dst, s_edges[1])
