"""Host time of the many-label path (slab.merge_slabs + host_graph.solve_periodic) on a noise grid, CPU only --
development aid.  The slabs' inputs are produced with the oracle and cached in /tmp.

    python tools/merge_profile.py [side] [p] [slabs] [--profile]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def inputs(side, p, n_slabs):
    path = f"/tmp/merge_profile_{side}_{p}_{n_slabs}.npz"
    if os.path.exists(path):
        z = np.load(path, allow_pickle=True)
        return [z[k] for k in ("nt", "nf", "boundary", "contacts", "bridges", "volumes")]
    import voxel_oracle as vo
    from test_slab_merge import plane_pairs_numpy
    from mdvcontainment_b200.slab import slab_extents
    grid = np.random.default_rng(5).random((side, side, side)) < p
    nt, nf, contacts, bridges, volumes, labels = [], [], [], [], [], []
    for x0, x1 in slab_extents(side, n_slabs):
        lab = vo.label_3d_grid(grid[x0:x1]).astype(np.int32)
        labels.append(lab)
        nt.append(int(max(lab.max(), 0))); nf.append(int(max(-lab.min(), 0)))
        contacts.append(vo.find_label_contacts_c(lab))
        b = vo.find_bridges_c(lab).reshape(-1, 5)
        bridges.append(b[b[:, 2] == 0])
        vol = np.zeros(nf[-1] + nt[-1] + 1, dtype=np.int64)
        v, c = np.unique(lab, return_counts=True)
        vol[v + nf[-1]] = c
        volumes.append(vol)
    boundary = [plane_pairs_numpy(labels[g][-1], labels[(g + 1) % n_slabs][0], 1 if g == n_slabs - 1 else 0)
                for g in range(n_slabs)]
    obj = lambda xs: np.array(xs + [None], dtype=object)[:-1]
    np.savez(path, nt=np.array(nt), nf=np.array(nf), boundary=obj(boundary), contacts=obj(contacts), bridges=obj(bridges),
             volumes=obj(volumes))
    return nt, nf, boundary, contacts, bridges, volumes


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    side = int(args[0]) if args else 256
    p = float(args[1]) if len(args) > 1 else 0.05
    n_slabs = int(args[2]) if len(args) > 2 else 4
    nt, nf, boundary, contacts, bridges, volumes = inputs(side, p, n_slabs)
    from mdvcontainment_b200 import host_graph as hg
    from mdvcontainment_b200.slab import merge_slabs

    def run():
        t0 = time.perf_counter()
        m = merge_slabs(nt, nf, boundary, contacts, bridges, volumes)
        t1 = time.perf_counter()
        NT, NF = m["n_true"], m["n_false"]
        uniq = np.concatenate([np.arange(-NF, 0, dtype=np.int32), np.arange(1, NT + 1, dtype=np.int32)])
        sol = hg.solve_periodic(uniq, m["contacts"], m["bridges"])
        lut = sol.lut(NF, NT)
        ids, vals = hg.counts_by_component(lut, m["volumes"])
        dag = sol.containment_dag(max_iterations=None)
        t2 = time.perf_counter()
        return m, sol, dag, t1 - t0, t2 - t1

    run()
    if "--profile" in sys.argv:
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        run()
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
    best = min((run()[3:] for _ in range(3)), key=sum)
    m, sol, dag, _, _ = run()
    print(f"{side}^3 p={p} {n_slabs} slabs: labels {m['n_true']} + {m['n_false']}, contacts {len(m['contacts'])}, "
          f"bridges {len(m['bridges'])}, components {len(sol.component_ids) if hasattr(sol, 'component_ids') else '?'}: "
          f"merge {best[0] * 1e3:.0f} ms, graph stage {best[1] * 1e3:.0f} ms")


if __name__ == "__main__":
    main()
