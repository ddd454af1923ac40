"""Host graph stage (mdvcontainment_b200/host_graph.py) against the goldens and the live reference."""
import numpy as np
import pytest

import voxel_oracle as vo
from golden_util import voxel_cases
from mdvcontainment_b200 import host_graph as hg

CASES = voxel_cases()


def solve(grid):
    labels = vo.label_3d_grid(grid).astype(np.int32)
    uniq = np.unique(labels)
    contacts, bridges = vo.find_label_contacts_c(labels), vo.find_bridges_c(labels)
    sol = hg.solve_periodic(uniq, contacts, bridges)
    comps = vo.create_components_grid(labels, sol["component2labels"])
    dag = hg.containment_graph(sol["is_contained"], np.unique(comps), sol["component_contact_graph"])
    counts = vo.counts(comps)
    return labels, sol, comps, dag, counts


@pytest.mark.parametrize("name", sorted(CASES))
def test_host_stage_matches_golden(name):
    c = CASES[name]
    if c["n_pos"] + c["n_neg"] > 4000:
        pytest.skip("many-label case covered by the reference comparison")
    labels, sol, comps, dag, counts = solve(c["grid"])
    assert [int(n) for n in dag.nodes] == c["nodes"]
    assert [[int(a), int(b)] for a, b in dag.edges()] == c["edges"]
    assert {str(int(k)): int(v) for k, v in sol["component_ranks"].items()} == c["ranks"]
    assert {str(int(k)): int(v) for k, v in counts.items()} == c["counts"]
    assert hg.format_dag_structure(dag, sol["component_ranks"], counts, unit="nvoxels") == c["str"]
    if c["has_arrays"]:
        assert np.array_equal(comps, c["components"])


@pytest.mark.reference
def test_host_stage_matches_live_reference_on_random_grids(ref):
    from mdvcontainment.containment_main import calc_containment_graph
    rng = np.random.default_rng(77)
    for k in range(150):
        hi = 11 if k < 120 else 19                      # the last ones have hundreds of labels
        shape = tuple(int(v) for v in rng.integers(1, hi, 3))
        g = rng.random(shape) < rng.choice([0.15, 0.35, 0.5, 0.65, 0.85] if k < 120 else [0.06, 0.15, 0.9])
        cg, ccg, comps_ref, ranks_ref, lcg = calc_containment_graph(g)
        labels, sol, comps, dag, counts = solve(g)
        assert np.array_equal(comps, comps_ref), shape
        assert list(dag.nodes) == list(cg.nodes) and list(dag.edges()) == list(cg.edges()), shape
        fast = sol.containment_dag()
        assert fast.nodes.tolist() == [int(n) for n in cg.nodes] and fast.edges() == [(int(u), int(v)) for u, v in cg.edges()]
        assert {int(a): int(b) for a, b in sol["component_ranks"].items()} == {int(a): int(b) for a, b in ranks_ref.items()}
        assert list(sol["component_ranks"]) == list(ranks_ref)
        assert list(sol["component_contact_graph"].nodes) == list(ccg.nodes)
        assert [list(ccg.neighbors(n)) for n in ccg.nodes] == [list(sol["component_contact_graph"].neighbors(n)) for n in ccg.nodes]
        nxg = sol["label_graph"].to_networkx()
        assert list(nxg.nodes) == list(lcg.nodes)
        assert [(u, v, str(d["label"])) for u, v, d in nxg.edges(data=True)] == [(u, v, str(d["label"])) for u, v, d in lcg.edges(data=True)]


def test_shift_graph_ranks_small_cases():
    """mdvc_host_shift_graph_ranks on hand-made graphs: rank of the lattice of closed-walk shift sums."""
    def ranks(n, edges):
        eu = [e[0] for e in edges]; ev = [e[1] for e in edges]; es = [e[2] for e in edges]
        root, rank, best = hg.shift_graph_ranks(n, eu, ev, es if edges else None)
        return root.tolist(), rank.tolist(), best
    assert ranks(3, []) == ([0, 1, 2], [0, 0, 0], 0)
    # a self bridge through z in both directions: rank 1
    assert ranks(1, [(0, 0, (0, 0, 1)), (0, 0, (0, 0, -1))])[1:] == ([1], 1)
    # two labels joined through a face and back inside the box: the loop 0 -> 1 -> 0 sums to (1, 0, 0)
    assert ranks(2, [(0, 1, (1, 0, 0)), (1, 0, (0, 0, 0))]) == ([0, 0], [1, 1], 1)
    # the same joined twice with opposite shifts only: a tree plus its reverse, no net displacement
    assert ranks(2, [(0, 1, (1, 0, 0)), (1, 0, (-1, 0, 0))]) == ([0, 0], [0, 0], 0)
    # three independent loops on one node, a dependent fourth
    loops = [(0, 0, (1, 1, 0)), (0, 0, (1, -1, 0)), (0, 0, (2, 0, 0))]
    assert ranks(1, loops)[2] == 2
    assert ranks(1, loops + [(0, 0, (1, 1, 1))])[2] == 3
    assert ranks(1, [(0, 0, (2, 4, 6)), (0, 0, (1, 2, 3))])[2] == 1
    # pieces are independent; skipped nodes cut the graph
    edges = [(0, 1, (0, 0, 0)), (1, 2, (0, 1, 0)), (2, 0, (0, 0, 0)), (3, 4, (0, 0, 0))]
    assert ranks(5, edges) == ([0, 0, 0, 3, 3], [1, 1, 1, 0, 0], 1)
    root, rank, best = hg.shift_graph_ranks(5, [e[0] for e in edges], [e[1] for e in edges], [e[2] for e in edges],
                                            skip=np.array([0, 1, 0, 0, 0], dtype=np.uint8))
    assert root.tolist() == [0, -1, 0, 3, 3] and rank.tolist() == [0, -1, 0, 0, 0] and best == 0
    # long chains keep exact potentials (path compression with offsets)
    n = 5000
    eu = list(range(n - 1)) + [n - 1]; ev = list(range(1, n)) + [0]
    es = [(1, 0, 0)] * (n - 1) + [(0, 0, 0)]
    assert hg.shift_graph_ranks(n, eu, ev, es)[2] == 1
    es[-1] = (-(n - 1), 0, 0)
    assert hg.shift_graph_ranks(n, eu, ev, es)[2] == 0


def test_view_order_is_what_networkx_iterates():
    """The numbering of the minority polarity follows the iteration order of a node-filtered networkx view."""
    import networkx as nx
    rng = np.random.default_rng(1)
    for n_neg, n_pos in [(1, 1), (3, 1), (1, 9), (5, 40), (40, 5), (300, 100), (7, 7), (2000, 30), (17, 5000), (0, 4), (4, 0)]:
        labels = np.concatenate([np.arange(-n_neg, 0), np.arange(1, n_pos + 1)]).astype(np.int32)
        g = nx.MultiDiGraph()
        g.add_nodes_from(int(v) for v in labels)
        for sub in (labels[labels > 0], labels[labels < 0]):
            want = [int(v) for v in g.subgraph(sub)]
            got = sub[hg._view_order(len(labels), sub.astype(np.int64))].tolist()
            assert got == want, (n_neg, n_pos)
    # arbitrary (non-contiguous) label values
    vals = np.unique(rng.integers(-5000, 5000, 600)).astype(np.int32)
    vals = vals[vals != 0]
    g = nx.MultiDiGraph()
    g.add_nodes_from(int(v) for v in vals)
    for sub in (vals[vals > 0][:50], vals[vals < 0][:37]):
        assert sub[hg._view_order(len(vals), sub.astype(np.int64))].tolist() == [int(v) for v in g.subgraph(sub)]


def _compare_with_model(grid):
    import graph_model as model
    labels = vo.label_3d_grid(grid).astype(np.int32)
    uniq = np.unique(labels)
    contacts, bridges = vo.find_label_contacts_c(labels), vo.find_bridges_c(labels)
    sol = hg.solve_periodic(uniq, contacts, bridges)
    ref = model.solve_periodic(uniq, contacts, bridges)
    assert {int(k): v for k, v in sol["component2labels"].items()} == {int(k): v for k, v in ref["component2labels"].items()}
    assert list(sol["component2labels"]) == list(ref["component2labels"])
    assert sol["component_ranks"] == ref["component_ranks"] and sol["complement_ranks"] == ref["complement_ranks"]
    assert sol["is_contained"] == ref["is_contained"]
    a, b = sol["component_contact_graph"], ref["component_contact_graph"]
    assert list(a.nodes) == list(b.nodes)
    assert [list(a.neighbors(n)) for n in a.nodes] == [list(b.neighbors(n)) for n in b.nodes]
    nf = int(max(-labels.min(), 0)); nt = int(max(labels.max(), 0))
    lut = sol.lut(nf, nt)
    for comp, labs in ref["component2labels"].items():
        assert (lut[np.asarray(labs) + nf] == comp).all()
    comps = np.unique(lut[lut != 0]) if (lut != 0).any() else np.array([], dtype=np.int32)
    want = model.containment_graph(ref["is_contained"], comps.astype(np.int32), ref["component_contact_graph"])
    dag = sol.containment_dag()
    assert dag.nodes.tolist() == [int(n) for n in want.nodes]
    assert dag.edges() == [(int(u), int(v)) for u, v in want.edges()]
    nxd = dag.to_networkx()
    assert list(nxd.nodes) == list(want.nodes) and list(nxd.edges()) == list(want.edges())
    counts = {int(c): 1 for c in comps}
    assert hg.format_dag_structure(dag, sol["component_ranks"], counts) == \
        model.format_dag_structure(want, ref["component_ranks"], counts)
    return len(uniq)


def test_array_solver_matches_the_plain_python_model_on_many_label_grids():
    rng = np.random.default_rng(5)
    seen = 0
    for shape, p in [((24, 24, 24), 0.05), ((24, 24, 24), 0.12), ((20, 26, 22), 0.3), ((16, 16, 40), 0.5),
                     ((30, 12, 18), 0.7), ((22, 22, 22), 0.93), ((9, 40, 11), 0.2), ((32, 32, 8), 0.08), ((48, 48, 48), 0.04)]:
        seen = max(seen, _compare_with_model(rng.random(shape) < p))
    assert seen > 2000                                 # the cases really have many labels
    from mdvcontainment_b200 import synth
    _compare_with_model(synth.salt_noise_grid((40, 24, 72), salt=0.05, seed=6))


def test_orientation_guard_and_level_sweep():
    """More than 100 000 components: the reference's orientation loop asserts; with the guard lifted (slab path) the
    level sweep gives a valid spanning forest of the same contact graph."""
    n = 100_500
    # one big positive label (1) touching n isolated negative labels, plus a negative background in contact with it
    contacts = np.stack([np.arange(-n - 1, 0), np.ones(n + 1, dtype=np.int64)], 1).astype(np.int32)
    bridges = np.array([[-(n + 1), -(n + 1), s, 0, 0] for s in (-1, 1)] + [[-(n + 1), -(n + 1), 0, s, 0] for s in (-1, 1)] +
                       [[-(n + 1), -(n + 1), 0, 0, s] for s in (-1, 1)], dtype=np.int32)
    uniq = np.concatenate([np.arange(-n - 1, 0), [1]]).astype(np.int32)
    sol = hg.solve_periodic(uniq, contacts, bridges)
    assert int(sol.comp_rank.max()) == 3 and int(sol.contained.sum()) == n + 1
    with pytest.raises(AssertionError, match="iterative death"):
        sol.containment_dag()
    dag = sol.containment_dag(max_iterations=None)
    assert len(dag) == n + 2 and len(dag.edge_src) == n + 1
    assert len(dag.roots()) == 1
    kids = dag.children()
    assert len(kids[dag.roots()[0]]) == 1 and len(kids[1]) == n


def test_many_label_case_matches_the_reference_recorded_graph_stage():
    """128^3 Bernoulli(0.05), 48 k labels: numbering, ranks, component contact graph and DAG
    as recorded from the unmodified reference (tests/golden/make_config_golden.py: 23.8 s there)."""
    import json
    import os
    import time
    from golden_util import GOLD, sha
    with open(os.path.join(GOLD, "config_cases.json")) as fh:
        rec = json.load(fh)["many_labels_128"]
    rng = np.random.default_rng(rec["seed"])
    grid = rng.random((rec["n"],) * 3) < rec["p"]
    assert sha(grid, np.uint8) == rec["sha_grid"]
    labels = vo.label_3d_grid(grid).astype(np.int32)
    contacts, bridges = vo.find_label_contacts_c(labels), vo.find_bridges_c(labels)
    assert sha(contacts) == rec["sha_contacts"] and sha(bridges) == rec["sha_bridges"]
    nf, nt = int(-labels.min()), int(labels.max())
    uniq = np.concatenate([np.arange(-nf, 0), np.arange(1, nt + 1)]).astype(np.int32)
    t0 = time.perf_counter()
    sol = hg.solve_periodic(uniq, contacts, bridges)
    lut = sol.lut(nf, nt)
    dag = sol.containment_dag()
    dt = time.perf_counter() - t0
    assert len(sol.comp_ids) == rec["n_components"]
    assert sha(lut) == rec["sha_lut"]
    assert sha(sol.comp_ids) == rec["sha_component_order"]
    assert sha(sol.comp_rank) == rec["sha_ranks"] and sha(sol.compl_rank) == rec["sha_complement_ranks"]
    assert sha(sol.contained.astype(np.uint8)) == rec["sha_contained"]
    ccg = sol["component_contact_graph"]
    assert sha(np.array([int(x) for x in ccg.nodes], dtype=np.int64)) == rec["sha_ccg_nodes"]
    assert sha(np.array([int(v) for x in ccg.nodes for v in ccg.neighbors(x)], dtype=np.int64)) == rec["sha_ccg_adjacency"]
    assert sha(dag.nodes.astype(np.int64)) == rec["dag"]["sha_nodes"]
    assert sha(np.array(dag.edges(), dtype=np.int64).reshape(-1, 2)) == rec["dag"]["sha_edges"]
    assert dt < 0.1 * rec["reference_graph_stage_seconds"]


def test_first_occurrence_scatter_path_equals_sort_path():
    """_first_occurrence with a key bound (two scatter passes) against the np.unique form, incl. the sizes where it
    falls back to the sort (bound far larger than the key count)."""
    rng = np.random.default_rng(3)
    for n, bound in ((1, 1), (7, 3), (1000, 40), (50000, 50000), (300, 10 ** 6), (0, 5)):
        keys = rng.integers(0, bound, n)
        assert np.array_equal(hg._first_occurrence(keys, bound), hg._first_occurrence(keys))


def test_adjacency_unique_equals_hash_version_on_symmetric_grouped_lists():
    """The component contact list is grouped by component, symmetric and free of self-pairs: keeping the rows with
    component < neighbour and skipping the de-duplication (mdvc_host_adjacency_unique) must give the adjacency lists
    the hash-table version builds from the whole list -- same neighbours in the same (networkx) order."""
    rng = np.random.default_rng(11)
    for n_comps, n_edges in ((2, 1), (5, 7), (60, 200), (3000, 9000)):
        a = rng.integers(0, n_comps, n_edges)
        b = rng.integers(0, n_comps, n_edges)
        keep = a != b
        pairs = np.unique(np.stack([np.minimum(a, b)[keep], np.maximum(a, b)[keep]], 1), axis=0)
        und = pairs[rng.permutation(len(pairs))]                       # undirected edges in a random "insertion" order
        src = np.concatenate([und[:, 0], und[:, 1]])
        dst = np.concatenate([und[:, 1], und[:, 0]])
        # grouped by component (ascending), each group's neighbours in some order of their own
        order = np.lexsort((rng.random(len(src)), src))
        cs, cd = src[order], dst[order]
        off_h, nb_h = hg.adjacency(n_comps, cs, cd)
        first = cs < cd
        off_u, nb_u = hg.adjacency(n_comps, cs[first], cd[first], unique=True)
        assert np.array_equal(off_h, off_u) and np.array_equal(nb_h, nb_u)
