"""BASELINE.json's full sizes on the CUDA path.

Config 3 (about 10.2 M beads binned into 1024^3, closing, both polarities): every stage against the known answers of
``tests/golden/fullsize_vesicle_1024.json`` -- produced without the CUDA path by oracle C binning, NumPy closing,
scipy.ndimage.label and the reference's compiled contact / bridge kernels (tests/golden/make_fullsize_golden.py) --
plus size-independent properties (closing is extensive and idempotent, equal-polarity neighbours share a label /
component, the components grid is the LUT image of the labels grid, the volumes add up).
Config 2 (1024^3 hollow cylinder): analytic voxel counts and the hierarchy of the 128^3 oracle run.
"""
import numpy as np
import pytest

import voxel_oracle as vo
from golden_util import fullsize_golden, sha, tensor_digest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

N = 1024
GOLD = fullsize_golden()


@pytest.fixture(scope="module")
def scene():
    from mdvcontainment_b200 import synth
    from mdvcontainment_b200.pipeline import FramePipeline
    if torch.cuda.mem_get_info()[0] < 48e9:
        pytest.skip("needs about 48 GB of free device memory")
    pos, dims = synth.vesicle_scene(box_nm=N / 2.0)
    pipe = FramePipeline(dims, 0.5, closing=True)
    out = pipe.run(torch.from_numpy(pos).cuda())
    torch.cuda.synchronize()
    yield {"pos": pos, "dims": dims, "pipe": pipe, "out": out}
    del pipe, out
    torch.cuda.empty_cache()


def test_scene_is_the_recorded_one(scene):
    if GOLD is None:
        pytest.skip("tests/golden/fullsize_vesicle_1024.json not generated")
    assert len(scene["pos"]) == GOLD["n_atoms"] and sha(scene["pos"]) == GOLD["sha_positions"]


def test_full_size_binning_matches_oracle(scene):
    """All 10.2 M beads: oracle C FMA chain on the host vs k_bin_atoms, voxel for voxel."""
    vox, nbox, _ = vo.voxelate_fma(scene["pos"], scene["dims"], 0.5)
    lin = (vox[:, 0] * N + vox[:, 1]) * N + vox[:, 2]
    want = torch.zeros(N ** 3, dtype=torch.bool, device="cuda")
    want[torch.from_numpy(lin).cuda()] = True
    got = scene["pipe"].bits_a.to_bool().reshape(-1)
    assert torch.equal(got, want)
    if GOLD is not None:
        assert int(got.sum()) == GOLD["n_occupied"] and tensor_digest(got) == GOLD["digest_occupancy"]


def test_full_size_closing(scene):
    from mdvcontainment_b200 import device as dev
    pipe = scene["pipe"]
    a, b = pipe.bits_a.words, pipe.bits_b.words
    assert not bool((a & ~b).any())                                   # extensive: closed >= binned
    again = dev.morph(pipe.bits_b, "de")
    assert torch.equal(again.words, b)                                # idempotent
    closed = pipe.bits_b.to_bool()
    if GOLD is not None:
        assert int(closed.sum()) == GOLD["n_closed"] and tensor_digest(closed) == GOLD["digest_closed"]
    # interiors of sub-blocks against the NumPy oracle (the 5x5x5 footprint of 'de' reaches 2 voxels)
    occ = pipe.bits_a.to_bool()
    for shift, (x0, y0, z0) in [((0, 0, 0), (464, 464, 200)), ((0, 0, 0), (470, 480, 800)), ((48, 48, 0), (0, 0, 208)),
                                ((40, 40, -216), (0, 0, 0))]:
        o = torch.roll(occ, shift, dims=(0, 1, 2))[x0:x0 + 96, y0:y0 + 96, z0:z0 + 96].cpu().numpy()
        c = torch.roll(closed, shift, dims=(0, 1, 2))[x0:x0 + 96, y0:y0 + 96, z0:z0 + 96].cpu().numpy()
        assert o.any()
        assert np.array_equal(vo.morph(o, "de")[2:-2, 2:-2, 2:-2], c[2:-2, 2:-2, 2:-2])


def _half_neighbourhood():
    return [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1) if (dx, dy, dz) > (0, 0, 0)]


def test_full_size_labels(scene):
    out, pipe = scene["out"], scene["pipe"]
    lab = out.labels_dev
    closed = pipe.bits_b.to_bool()
    assert torch.equal(lab > 0, closed) and not bool((lab == 0).any())
    assert int(lab.max()) == out.n_true and int(lab.min()) == -out.n_false
    # soundness: two 26-neighbours (no wrap) of equal polarity carry the same label
    for dx, dy, dz in _half_neighbourhood():
        def cut(t, sign):
            sl = []
            for d, n in zip((dx, dy, dz), t.shape):
                d *= sign
                sl.append(slice(max(d, 0), n + min(d, 0)))
            return t[tuple(sl)]
        a, b = cut(lab, 1), cut(lab, -1)
        assert not bool((((a > 0) == (b > 0)) & (a != b)).any()), (dx, dy, dz)
    if GOLD is not None:
        assert (out.n_true, out.n_false) == (GOLD["n_true"], GOLD["n_false"])
        assert tensor_digest(lab) == GOLD["digest_labels"]            # scipy's numbering, voxel for voxel
        vol = pipe.plan.label_volumes()
        assert [int(v) for v in vol[:len(GOLD["label_volumes"])]] == GOLD["label_volumes"]


def test_full_size_contacts_and_bridges(scene):
    if GOLD is None:
        pytest.skip("tests/golden/fullsize_vesicle_1024.json not generated")
    out = scene["out"]
    assert np.asarray(out.contacts).reshape(-1, 2).tolist() == GOLD["contacts"]
    assert np.asarray(out.bridges).reshape(-1, 5).tolist() == GOLD["bridges"]


def test_full_size_components(scene):
    out, pipe = scene["out"], scene["pipe"]
    comp, lab = out.components_dev, out.labels_dev
    lut = pipe.plan.lut()                                             # index = label + n_false
    lut_dev = torch.as_tensor(np.asarray(lut), device="cuda")
    for x0 in range(0, N, 128):                                       # components == LUT[labels], slab by slab
        assert torch.equal(comp[x0:x0 + 128], lut_dev[(lab[x0:x0 + 128] + out.n_false).long()])
    # periodic soundness: equal-polarity 26-neighbours across the box faces share a component
    closed = pipe.bits_b.to_bool()
    for d in _half_neighbourhood():
        same = torch.roll(closed, d, dims=(0, 1, 2)) == closed
        assert not bool((same & (torch.roll(comp, d, dims=(0, 1, 2)) != comp)).any()), d
    counts = {int(k): int(v) for k, v in out.voxel_counts.items()}
    assert sum(counts.values()) == N ** 3
    ids, cnt = torch.unique(comp, return_counts=True)
    assert dict(zip(ids.tolist(), cnt.tolist())) == counts
    # the scene: slab and two nested vesicles in solvent -> 3 solid and 3 solvent components; the hierarchy is the
    # one the same scene gives at 192^3 (a size the golden-pinned small-grid tests cover)
    assert sorted(k > 0 for k in counts) == [False] * 3 + [True] * 3
    import networkx as nx
    from mdvcontainment_b200 import synth
    from mdvcontainment_b200.pipeline import FramePipeline
    pos_s, dims_s = synth.vesicle_scene(box_nm=96.0)
    small = FramePipeline(dims_s, 0.5, closing=True).run(torch.from_numpy(pos_s).cuda())

    def tagged(o):
        g = nx.DiGraph(o.containment_graph)
        nx.set_node_attributes(g, {n: (int(o.component_ranks[n]), int(n) > 0) for n in g.nodes}, "sig")
        return g
    assert nx.is_isomorphic(tagged(out), tagged(small), node_match=lambda a, b: a["sig"] == b["sig"])


def test_hollow_cylinder_1024_known_answer():
    """Config 2: counts are nz times the 2-D cross-section's; the hierarchy is the one of the 128^3 oracle run."""
    from mdvcontainment_b200 import VoxelContainment, device as dev, synth
    ring2d = synth.hollow_cylinder_grid(N, 1)[:, :, 0]
    bits = dev.BitGrid.from_bool(torch.from_numpy(ring2d).cuda()[:, :, None].expand(N, N, N).contiguous())
    vc = VoxelContainment(bits)
    r = np.hypot(*np.meshgrid(np.arange(N) - N / 2 + 0.5, np.arange(N) - N / 2 + 0.5, indexing="ij"))
    inside, ring, outside = int((r < 0.2 * N).sum()), int(ring2d.sum()), int((r >= 0.3 * N).sum())
    counts = {int(k): int(v) for k, v in vc.voxel_counts.items()}
    assert sorted(counts.values()) == sorted([inside * N, ring * N, outside * N])
    small = VoxelContainment(synth.hollow_cylinder_grid(128))
    assert sorted(vc.component_ranks.values()) == sorted(small.component_ranks.values())
    assert [sorted(len(list(v.containment_graph.successors(n))) for n in v.containment_graph.nodes) for v in (vc, small)][0] == \
        sorted(len(list(small.containment_graph.successors(n))) for n in small.containment_graph.nodes)
    assert str(vc).count("\n") == str(small).count("\n")
    # the same grid through the unmodified reference (tests/golden/make_config_golden.py --only cylinder1024): DAG text,
    # contacts, bridges and digests of the label and component grids
    import json
    import os
    from golden_util import GOLD
    with open(os.path.join(GOLD, "config_cases.json")) as fh:
        rec = json.load(fh).get("cylinder_1024")
    if rec is not None:
        closed = dev.morph(bits, "de")
        assert tensor_digest(closed.to_bool()) == rec["digest_closed"]
        ref_vc = VoxelContainment(closed)
        assert str(ref_vc) == rec["str"]
        assert {str(int(k)): int(v) for k, v in ref_vc.voxel_counts.items()} == rec["counts"]
        assert np.asarray(ref_vc._frame.contacts).reshape(-1, 2).tolist() == rec["contacts"]
        assert np.asarray(ref_vc._frame.bridges).reshape(-1, 5).tolist() == rec["bridges"]
        assert tensor_digest(ref_vc.nonp_labeled_grid_device) == rec["digest_labels"]
        assert tensor_digest(ref_vc.components_grid_device) == rec["digest_components"]
        del ref_vc, closed
    del vc, bits
    torch.cuda.empty_cache()
