"""The run-level rules used by csrc/frame.cu, validated on the CPU against the oracle/goldens."""
import numpy as np
import pytest

import run_model as rm
import voxel_oracle as vo
from golden_util import voxel_cases

CASES = voxel_cases()
SMALL = sorted(n for n, c in CASES.items() if np.prod(c["shape"]) <= 12000)


@pytest.mark.parametrize("name", SMALL)
def test_run_model_matches_golden(name):
    c = CASES[name]
    g = c["grid"]
    runstart, rowoff, lab, nt, nf = rm.label_runs(g)
    assert (nt, nf) == (c["n_pos"], c["n_neg"])
    labels = rm.expand(g, runstart, lab)
    assert np.array_equal(labels, c["labels"])
    contacts, bridges = rm.pairs(g, runstart, rowoff, lab)
    assert np.array_equal(contacts, c["contacts"].reshape(-1, 2))
    assert np.array_equal(bridges, c["bridges"])
    # the device-side numbering rule yields the reference's partition (numbering itself is the host's)
    lut, kp, kn = rm.component_lut(bridges, nt, nf)
    comps = lut[labels + nf]
    pairs = np.unique(np.stack([comps.reshape(-1), c["components"].reshape(-1)], 1), axis=0)
    assert len(pairs) == kp + kn == len(np.unique(c["components"]))


def test_run_model_random_shapes():
    rng = np.random.default_rng(123)
    for _ in range(150):
        shape = tuple(int(v) for v in rng.integers(1, 8, 3))
        g = rng.random(shape) < rng.choice([0.2, 0.5, 0.8])
        runstart, rowoff, lab, nt, nf = rm.label_runs(g)
        want = vo.label_3d_grid(g).astype(np.int32)
        assert np.array_equal(rm.expand(g, runstart, lab), want)
        contacts, bridges = rm.pairs(g, runstart, rowoff, lab)
        assert np.array_equal(contacts, vo.find_label_contacts_c(want))
        assert np.array_equal(bridges, vo.find_bridges_c(want))


def _smooth_grids(rng, n):
    """Blobby grids (few long runs per row, many mirror rows) -- where the shortcuts actually fire."""
    for _ in range(n):
        shape = tuple(int(v) for v in rng.integers(3, 12, 3))
        f = rng.random(shape)
        for ax in range(3):
            f = f + np.roll(f, 1, ax) + np.roll(f, -1, ax)
        yield f > np.quantile(f, rng.choice([0.3, 0.5, 0.7]))


def test_shortcut_rules_give_the_same_answer():
    """Mirror rows (positional link partner, no cross-row contacts) and contacts from voxel-on-voxel overlaps only:
    the three work-skipping rules of the link / pair kernels, against the oracle on random and on smooth grids."""
    rng = np.random.default_rng(321)
    grids = [rng.random(tuple(int(v) for v in rng.integers(1, 8, 3))) < rng.choice([0.2, 0.5, 0.8]) for _ in range(120)]
    grids += list(_smooth_grids(rng, 60))
    fired = 0
    for g in grids:
        runstart, rowoff, lab, nt, nf = rm.label_runs(g, shortcuts=True)
        want = vo.label_3d_grid(g).astype(np.int32)
        assert np.array_equal(rm.expand(g, runstart, lab), want), g.shape
        contacts, bridges = rm.pairs(g, runstart, rowoff, lab, shortcuts=True)
        assert np.array_equal(contacts, vo.find_label_contacts_c(want)), g.shape
        assert np.array_equal(bridges, vo.find_bridges_c(want)), g.shape
        X, Y, Z = g.shape
        flat = g.reshape(-1)
        fired += sum(rm.rows_mirror(runstart, rowoff, flat, Z, r, r - 1) for r in range(1, X * Y)
                     if rowoff[r + 1] - rowoff[r] > 1)
    assert fired > 100          # the shortcut was exercised on rows with more than one run
