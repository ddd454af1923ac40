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
    for k in range(120):
        shape = tuple(int(v) for v in rng.integers(1, 11, 3))
        g = rng.random(shape) < rng.choice([0.15, 0.35, 0.5, 0.65, 0.85])
        cg, ccg, comps_ref, ranks_ref, lcg = calc_containment_graph(g)
        labels, sol, comps, dag, counts = solve(g)
        assert np.array_equal(comps, comps_ref), shape
        assert list(dag.nodes) == list(cg.nodes) and list(dag.edges()) == list(cg.edges()), shape
        assert {int(a): int(b) for a, b in sol["component_ranks"].items()} == {int(a): int(b) for a, b in ranks_ref.items()}
        assert list(sol["component_ranks"]) == list(ranks_ref)
        assert list(sol["component_contact_graph"].nodes) == list(ccg.nodes)
        assert [list(ccg.neighbors(n)) for n in ccg.nodes] == [list(sol["component_contact_graph"].neighbors(n)) for n in ccg.nodes]
        nxg = sol["label_graph"].to_networkx()
        assert list(nxg.nodes) == list(lcg.nodes)
        assert [(u, v, str(d["label"])) for u, v, d in nxg.edges(data=True)] == [(u, v, str(d["label"])) for u, v, d in lcg.edges(data=True)]


def test_integer_rank():
    assert hg._int_rank([]) == 0
    assert hg._int_rank([(0, 0, 1), (0, 0, -1)]) == 1
    assert hg._int_rank([(1, 1, 0), (1, -1, 0), (2, 0, 0)]) == 2
    assert hg._int_rank([(1, 0, 0), (0, 1, 0), (1, 1, 1)]) == 3
    assert hg._int_rank([(2, 4, 6), (1, 2, 3)]) == 1
