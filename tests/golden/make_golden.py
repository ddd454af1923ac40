"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Build container only (needs /root/reference):   python tests/golden/make_golden.py
Outputs: voxel_cases.npz, binning_cases.npz, e2e_cases.json  (small; committed).
The reference package is imported through oracle/refshim.py (matplotlib / MDAnalysis stubs,
Cython kernels compiled by oracle/build_ref.py from the sources where they lie).
"""
import contextlib
import hashlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import refshim  # noqa: E402

ref = refshim.load_reference()
from mdvcontainment import Containment, VoxelContainment  # noqa: E402
from mdvcontainment.atomgroup_to_voxels import create_voxels, morph_voxels  # noqa: E402
from mdvcontainment.containment_main import calc_containment_graph  # noqa: E402
from mdvcontainment.find_bridges import find_bridges  # noqa: E402
from mdvcontainment.find_label_contacts import find_label_contacts  # noqa: E402
from mdvcontainment.gen_data import create_3d_boolean_grid, make_complex_3D  # noqa: E402
from mdvcontainment.label import label_3d_grid  # noqa: E402

from mdvcontainment_b200 import synth  # noqa: E402
from mdvcontainment_b200.universe import Universe, read_gro, voxels_to_universe  # noqa: E402


def sha(a, dtype=None):
    a = np.ascontiguousarray(a if dtype is None else np.asarray(a).astype(dtype))
    return hashlib.sha256(a.tobytes()).hexdigest()[:16]


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = fn(*a, **k)
    return out, buf.getvalue()


# ---------------------------------------------------------------------------
def voxel_cases():
    cases = {}
    cx = {r: make_complex_3D(r) for r in (0, 3, 11)}
    for r, g in cx.items():
        cases[f"complex3D_roll{r}"] = g
    cases["complex3D_roll11_closed"] = morph_voxels(cx[11], "de")
    perlin = [((32, 32, 32), 0.5, [2, 2, 2], 0), ((64, 64, 64), 0.5, [4, 4, 4], 1),
              ((48, 32, 16), 0.45, [4, 4, 4], 2), ((64, 64, 64), 0.55, [8, 8, 8], 3)]
    for shape, p, res, seed in perlin:
        g = create_3d_boolean_grid(shape, p, res, seed)
        name = "perlin_%dx%dx%d_s%d" % (*shape, seed)
        cases[name] = g
        cases[name + "_closed"] = morph_voxels(g, "de")
    rng = np.random.default_rng(42)
    for shape in [(1, 1, 1), (1, 1, 5), (2, 3, 70), (5, 1, 33), (3, 3, 3), (7, 70, 3), (33, 34, 35),
                  (2, 2, 2), (1, 40, 1), (17, 5, 64), (9, 9, 96), (4, 4, 31), (6, 5, 32), (12, 11, 65)]:
        for p in (0.15, 0.5, 0.85):
            cases["rand_%dx%dx%d_p%02d" % (*shape, int(p * 100))] = rng.random(shape) < p
    cases["allfalse_4x5x6"] = np.zeros((4, 5, 6), bool)
    cases["alltrue_4x5x6"] = np.ones((4, 5, 6), bool)
    cases["cylinder64_closed"] = morph_voxels(synth.hollow_cylinder_grid(64), "de")
    cases["salt48"] = synth.salt_noise_grid((48, 48, 48), seed=5)
    cases["salt_40x24x72"] = synth.salt_noise_grid((40, 24, 72), salt=0.05, seed=6)
    return cases


def dump_voxel_cases():
    store, meta = {}, {}
    for name, g in voxel_cases().items():
        g = np.ascontiguousarray(g, dtype=bool)
        labels = label_3d_grid(g).astype(np.int32)
        contacts = find_label_contacts(labels)
        bridges = find_bridges(labels)
        (cg, ccg, comps, ranks, lcg), _ = quiet(calc_containment_graph, g)
        vc, _ = quiet(VoxelContainment, g)
        store[name + "/grid_bits"] = np.packbits(g.reshape(-1))
        store[name + "/contacts"] = contacts
        store[name + "/bridges"] = bridges
        small = g.size <= 40000
        if small:
            store[name + "/labels"] = labels
            store[name + "/components"] = comps.astype(np.int32)
        lab_vals, lab_cnt = np.unique(labels, return_counts=True)
        meta[name] = {
            "shape": list(g.shape), "n_true": int(g.sum()), "sha_grid": sha(g, np.uint8),
            "n_pos": int(max(labels.max(), 0)), "n_neg": int(max(-labels.min(), 0)),
            "sha_labels": sha(labels), "sha_contacts": sha(contacts), "sha_bridges": sha(bridges),
            "sha_components": sha(comps.astype(np.int32)),
            "label_volumes": {str(int(k)): int(v) for k, v in zip(lab_vals, lab_cnt)} if len(lab_vals) <= 64 else None,
            "counts": {str(int(k)): int(v) for k, v in vc.voxel_counts.items()},
            "ranks": {str(int(k)): int(v) for k, v in vc.component_ranks.items()},
            "nodes": [int(n) for n in vc.containment_graph.nodes],
            "edges": [[int(a), int(b)] for a, b in vc.containment_graph.edges()],
            "str": str(vc), "has_arrays": small,
        }
    np.savez_compressed(os.path.join(HERE, "voxel_cases.npz"), **store)
    with open(os.path.join(HERE, "voxel_cases.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    return meta


# ---------------------------------------------------------------------------
def binning_inputs():
    """name -> (positions float32 [n,3] A, dimensions float32[6], resolution nm)."""
    out = {}
    rng = np.random.default_rng(11)
    boxes = {"ortho": (100.0, 80.0, 60.0, 90, 90, 90), "cubic_odd": (97.3, 97.3, 45.1, 90, 90, 90),
             "hex": (120.0, 120.0, 90.0, 90, 90, 60), "rhombic": (100.0, 100.0, 100.0, 60, 60, 90),
             "triclinic": (64.5, 71.25, 80.0, 75, 80, 85)}
    for bname, dims in boxes.items():
        for res in (0.5, 1.0, 0.37):
            n = 1500
            d32 = np.array(dims, dtype=np.float32)
            pos = (rng.random((n, 3)) * 3 - 1) * d32[:3]
            pos[: n // 2] = np.round(pos[: n // 2] / (10 * res)) * (10 * res)
            out[f"{bname}_r{res}"] = (np.round(pos, 2).astype(np.float32), d32, res)
    return out


def dump_binning_cases():
    store = {}
    for name, (pos, dims, res) in binning_inputs().items():
        u = Universe(pos.copy(), dims)
        store[name + "/positions"] = pos
        store[name + "/dimensions"] = dims
        store[name + "/resolution"] = np.float64(res)
        for p in (1, 2, 3):   # three successive passes: the position write-back is not idempotent
            grid, mapping = create_voxels(u.atoms, res, return_mapping=True)
            store[f"{name}/pass{p}_voxels"] = mapping["atom_voxels"].astype(np.int32)
            store[f"{name}/pass{p}_positions"] = u.atoms.positions.copy()
            store[f"{name}/pass{p}_grid_bits"] = np.packbits(grid.reshape(-1))
            store[f"{name}/grid_shape"] = np.array(grid.shape)
    np.savez_compressed(os.path.join(HERE, "binning_cases.npz"), **store)


# ---------------------------------------------------------------------------
def e2e_universes():
    """name -> (builder returning (universe, selection), kwargs list); inputs shared with the tests."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import e2e_inputs

    def complex3d():
        u, sel = e2e_inputs.build("complex3D", make_complex_3D(11))
        gro = os.path.join(refshim.REF_ROOT, "examples/structures/complex3D.gro")
        ug = read_gro(gro)                       # the bundled fixture is exactly this universe
        assert np.array_equal(ug.atoms.positions, u.atoms.positions)
        assert list(ug.atoms.names) == list(u.atoms.names)
        assert np.array_equal(ug.dimensions, u.dimensions)
        return u, sel

    out = {}
    for name, runs in e2e_inputs.RUNS.items():
        out[name] = (complex3d if name == "complex3D" else (lambda n=name: e2e_inputs.build(n)), runs)
    return out


def dump_e2e_cases():
    out = {}
    for name, (build, runs) in e2e_universes().items():
        for kw in runs:
            u, sel = build()
            in_sha = sha(u.atoms.positions)
            c, printed = quiet(Containment, sel, **kw)
            vc = c.voxel_containment
            rec = {
                "kwargs": kw, "n_atoms": len(u.atoms), "n_selected": len(sel), "sha_input_positions": in_sha,
                "printed": printed, "str": str(c), "str_voxels": str(vc),
                "grid_shape": list(c.boolean_grid.shape), "sha_grid": sha(c.boolean_grid, np.uint8),
                "sha_components": sha(vc.components_grid.astype(np.int32)),
                "sha_positions_after": sha(u.atoms.positions),
                "voxel_volume": float(c.voxel_volume),
                "nodes": [int(n) for n in vc.nodes],
                "edges": [[int(a), int(b)] for a, b in vc.containment_graph.edges()],
                "ranks": {str(int(k)): int(v) for k, v in vc.component_ranks.items()},
                "counts": {str(int(k)): int(v) for k, v in vc.voxel_counts.items()},
            }
            if not kw.get("no_mapping", False):
                tf = u.atoms.tempfactors
                vals, cnt = np.unique(tf, return_counts=True)
                rec["sha_tempfactors"] = sha(tf, np.float64)
                rec["tempfactor_hist"] = {str(int(v)): int(n) for v, n in zip(vals, cnt)}
                per_node = {}
                for node in vc.nodes:
                    a = c.get_atomgroup_from_nodes([node])
                    b = c.get_atomgroup_from_nodes([node], containment=True)
                    per_node[str(int(node))] = [len(a), sha(a.ix), len(b), sha(b.ix)]
                rec["atoms_per_node"] = per_node
                if not kw.get("slab", False):
                    big = max(vc.voxel_counts.values()) // 20 + 1
                    view = c.node_view(min_size=int(big))
                    rec["view_min_size"] = int(big)
                    rec["view_str"] = str(view)
                    rec["view_nodes"] = [int(n) for n in view.nodes]
                    rec["view_atoms"] = {str(int(n)): [len(view.get_atomgroup_from_nodes([n])),
                                                       sha(view.get_atomgroup_from_nodes([n]).ix)]
                                         for n in view.nodes}
            key = name + "|" + json.dumps(kw, sort_keys=True)
            out[key] = rec
            print(key, "->", rec["str"].splitlines()[0])
    with open(os.path.join(HERE, "e2e_cases.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    m = dump_voxel_cases()
    print(f"voxel cases: {len(m)}")
    dump_binning_cases()
    dump_e2e_cases()
