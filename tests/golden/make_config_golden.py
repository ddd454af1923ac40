"""Golden answers for BASELINE.json's configs 2, 4 and the many-label host-graph case, produced by the UNMODIFIED
reference (build container only; needs /root/reference):

    python tests/golden/make_config_golden.py            -> tests/golden/config_cases.json

* config 2: gen_data-style hollow cylinder at 512^3 (generator anchor gen_data.py:13-14), closing
  'de', through the reference's ``morph_voxels`` + ``VoxelContainment``: DAG text, counts, contacts, bridges and
  64-bit digests of the closed grid, the label grid and the components grid (about 20 s, 3 GB).
* config 4: frames of ``synth.trajectory_frame`` at a 32 nm box (64^3 voxels) through the reference's
  ``Containment(selection, 0.5, closing=True, no_mapping=True, betafactors=False)``; an NVT set (one box) and an NPT
  set (box breathing by +-1.2 %, which changes the transform every frame and the grid shape for some).
* many labels: 128^3 Bernoulli(0.05) (48 k labels).  The reference's
  ``create_components_grid`` needs 90 s and its graph stage 23 s there, so only the graph stage is run -- contact
  graph, subgraphs, relabel dicts, ranks, containment flags, component contact graph, containment graph
  (containment_main.py:72-133 without the grid passes) -- and recorded as hashes.
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(HERE)]

import refshim  # noqa: E402

ref = refshim.load_reference()
from mdvcontainment import Containment, VoxelContainment  # noqa: E402
from mdvcontainment.atomgroup_to_voxels import morph_voxels  # noqa: E402
from mdvcontainment.find_bridges import find_bridges  # noqa: E402
from mdvcontainment.find_label_contacts import find_label_contacts  # noqa: E402
from mdvcontainment.label import label_3d_grid  # noqa: E402
from mdvcontainment import graph_logic as gl, rank_logic as rl  # noqa: E402

from e2e_inputs import npt_dimensions  # noqa: E402
from golden_util import sha  # noqa: E402
from mdvcontainment_b200 import synth  # noqa: E402
from mdvcontainment_b200.universe import Universe  # noqa: E402


def digest(a):
    """golden_util.tensor_digest in NumPy: sum(value[i] * (i mod 1000003 + 1)) in wrapping int64."""
    flat = np.ascontiguousarray(a).reshape(-1)
    acc = np.int64(0)
    with np.errstate(over="ignore"):
        for s in range(0, flat.size, 1 << 26):
            part = flat[s:s + (1 << 26)].astype(np.int64)
            w = np.arange(s, s + part.size, dtype=np.int64) % 1000003 + 1
            acc = np.int64(acc + (part * w).sum(dtype=np.int64))
    return int(acc)


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        return fn(*a, **k)


def cylinder_512(n=512):
    t0 = time.time()
    grid = synth.hollow_cylinder_grid(n)
    closed = morph_voxels(grid, "de")
    vc = quiet(VoxelContainment, closed)
    labels = label_3d_grid(closed).astype(np.int32)
    rec = {
        "n": n, "str": str(vc), "seconds": time.time() - t0,
        "counts": {str(int(k)): int(v) for k, v in vc.voxel_counts.items()},
        "ranks": {str(int(k)): int(v) for k, v in vc.component_ranks.items()},
        "nodes": [int(x) for x in vc.nodes],
        "edges": [[int(a), int(b)] for a, b in vc.containment_graph.edges()],
        "contacts": find_label_contacts(labels).tolist(), "bridges": find_bridges(labels).tolist(),
        "n_true_input": int(grid.sum()), "n_true_closed": int(closed.sum()),
        "digest_closed": digest(closed), "digest_labels": digest(labels),
        "digest_components": digest(vc.components_grid.astype(np.int32)),
    }
    return rec


def trajectory(frames, box_nm=32.0, npt=False):
    out = {}
    for f in frames:
        pos, dims = synth.trajectory_frame(f, box_nm=box_nm)
        if npt:
            dims = npt_dimensions(f, box_nm)
        u = Universe(pos.copy(), dims)
        c = quiet(Containment, u.atoms, 0.5, closing=True, no_mapping=True, betafactors=False)
        vc = c.voxel_containment
        out[str(f)] = {
            "grid_shape": list(c.boolean_grid.shape), "sha_grid": sha(c.boolean_grid, np.uint8),
            "sha_components": sha(vc.components_grid.astype(np.int32)), "str": str(vc),
            "nodes": [int(x) for x in vc.nodes], "edges": [[int(a), int(b)] for a, b in vc.containment_graph.edges()],
            "ranks": {str(int(k)): int(v) for k, v in vc.component_ranks.items()},
            "counts": {str(int(k)): int(v) for k, v in vc.voxel_counts.items()},
        }
        print("frame", f, "npt" if npt else "nvt", out[str(f)]["grid_shape"], out[str(f)]["str"].splitlines()[0])
    return out


def many_labels(n=128, p=0.05, seed=0):
    rng = np.random.default_rng(seed)
    grid = rng.random((n, n, n)) < p
    labels = label_3d_grid(grid)
    uniq = np.unique(labels)
    contacts, bridges = find_label_contacts(labels), find_bridges(labels)
    t0 = time.time()
    cg = gl.create_contact_graph(contacts, uniq, bridges)
    pos_sub, neg_sub = gl.get_subgraphs(uniq, cg)
    c2l, ls2c, l2c = gl.get_mapping_dicts(pos_sub, neg_sub)
    comp_ranks, compl_ranks = rl.get_ranks(pos_sub, neg_sub, uniq, c2l, ls2c, cg)
    contained = rl.get_is_contained(comp_ranks, compl_ranks)
    ccg = gl.create_component_contact_graph(c2l, l2c, cg)
    # components_grid is not built (90 s): np.unique(components_grid) is the sorted component ids
    unique_components = np.array(sorted(int(c) for c in c2l), dtype=np.int32)
    t1 = time.time()
    try:
        dag = gl.create_containment_graph(contained, unique_components, ccg)
        dag_rec = {"sha_nodes": sha(np.array([int(x) for x in dag.nodes], dtype=np.int64)),
                   "sha_edges": sha(np.array([[int(a), int(b)] for a, b in dag.edges()], dtype=np.int64).reshape(-1, 2)),
                   "n_edges": dag.number_of_edges()}
    except AssertionError as exc:                      # more than 100 000 components
        dag_rec = {"assertion": str(exc)}
    nf, nt = int(max(-labels.min(), 0)), int(max(labels.max(), 0))
    lut = np.zeros(nf + nt + 1, dtype=np.int32)
    for comp, labs in c2l.items():
        lut[np.asarray(labs, dtype=np.int64) + nf] = comp
    comp_order = [int(c) for c in c2l]
    rec = {
        "n": n, "p": p, "seed": seed, "sha_grid": sha(grid, np.uint8), "n_labels": int(len(uniq)),
        "n_contacts": int(len(contacts)), "n_bridges": int(len(bridges)),
        "sha_contacts": sha(contacts), "sha_bridges": sha(bridges),
        "n_components": len(c2l), "sha_lut": sha(lut),
        "sha_component_order": sha(np.array(comp_order, dtype=np.int64)),
        "sha_ranks": sha(np.array([int(comp_ranks[c]) for c in c2l], dtype=np.int64)),
        "sha_complement_ranks": sha(np.array([int(compl_ranks[c]) for c in c2l], dtype=np.int64)),
        "sha_contained": sha(np.array([bool(contained[c]) for c in c2l], dtype=np.uint8)),
        "sha_ccg_nodes": sha(np.array([int(x) for x in ccg.nodes], dtype=np.int64)),
        "sha_ccg_adjacency": sha(np.array([int(v) for x in ccg.nodes for v in ccg.neighbors(x)], dtype=np.int64)),
        "reference_graph_stage_seconds": t1 - t0, "dag": dag_rec,
    }
    print("many labels:", rec["n_labels"], "labels,", rec["n_components"], "components, reference graph stage",
          round(t1 - t0, 1), "s")
    return rec


def main():
    out_path = os.path.join(HERE, "config_cases.json")
    rec = {}
    if os.path.exists(out_path) and "--only" in sys.argv:
        with open(out_path) as fh:
            rec = json.load(fh)
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    frames = [0, 1, 2, 3, 5, 17, 33, 50, 75, 99]
    if only in (None, "trajectory"):
        rec["trajectory_nvt"] = {"box_nm": 32.0, "frames": trajectory(frames)}
        rec["trajectory_npt"] = {"box_nm": 32.0, "frames": trajectory(frames, npt=True)}
    if only in (None, "many"):
        rec["many_labels_128"] = many_labels()
    if only in (None, "cylinder"):
        rec["cylinder_512"] = cylinder_512()
        print("cylinder 512:", rec["cylinder_512"]["str"])
    if only == "cylinder1024":                         # ~10 min and ~25 GB of host memory: run on request only
        rec["cylinder_1024"] = cylinder_512(1024)
        print("cylinder 1024:", rec["cylinder_1024"]["str"])
    with open(out_path, "w") as fh:
        json.dump(rec, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
