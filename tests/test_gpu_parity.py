"""Parity of the CUDA path (through the C-ABI) with the oracle and the golden vectors. Needs a GPU."""
import numpy as np
import pytest

import run_model as rm
import voxel_oracle as vo
from golden_util import binning_cases, sha, voxel_cases

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

CASES = voxel_cases()


@pytest.fixture(scope="module")
def dev():
    from mdvcontainment_b200 import device
    return device


def rand_grids(n, seed, maxdim=12, big_z=True):
    rng = np.random.default_rng(seed)
    for i in range(n):
        shape = [int(v) for v in rng.integers(1, maxdim, 3)]
        if big_z and i % 3 == 0:
            shape[2] = int(rng.integers(30, 140))
        yield rng.random(shape) < rng.choice([0.1, 0.3, 0.5, 0.7, 0.9])


# ------------------------------------------------------------------------------------------ T1
def test_pack_unpack_popcount(dev):
    for g in rand_grids(40, 1):
        b = dev.BitGrid.from_bool(g)
        assert np.array_equal(b.to_bool().cpu().numpy(), g)
        assert b.count() == int(g.sum())
        words = b.words.cpu().numpy().view(np.uint32)
        Z = g.shape[2]
        # tail bits and pad words are zero
        for w in range(b.pitch):
            nbits = min(max(Z - 32 * w, 0), 32)
            mask = np.uint32(0xFFFFFFFF if nbits == 32 else (1 << nbits) - 1)
            assert not (words[:, :, w] & ~mask).any()


def test_to_numpy_staged_copy(dev, monkeypatch):
    """Large device tensors reach NumPy through two pinned staging buffers; same bytes as .cpu().numpy()."""
    monkeypatch.setattr(dev, "_PINNED_BYTES", 1 << 20)
    monkeypatch.setattr(dev, "_PINNED", [])
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randint(-2 ** 31, 2 ** 31 - 1, (130, 77, 301), dtype=torch.int32, device="cuda", generator=g)   # 12 MB, ragged last chunk
    for t in (a, a > 0, a.to(torch.int64), a.float(), a[:, ::2, :], a.permute(2, 0, 1)):
        got = dev.to_numpy(t)
        want = t.cpu().numpy()
        assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want)
    small = a[:2]
    assert np.array_equal(dev.to_numpy(small), small.cpu().numpy())
    for h in (a.cpu().numpy(), (a > 0).cpu().numpy(), a.float().cpu().numpy(), a.cpu().numpy()[:, ::3, :]):
        d = dev.to_device(h)
        assert d.is_cuda and d.shape == h.shape and np.array_equal(d.cpu().numpy(), h)


# ------------------------------------------------------------------------------------------ A3/A4
@pytest.mark.parametrize("ops", ["d", "e", "de", "ed", "dde", "deed", ""])
def test_morph_matches_oracle(dev, ops):
    for g in rand_grids(25, 2):
        got = dev.morph(dev.BitGrid.from_bool(g), ops).to_bool().cpu().numpy()
        assert np.array_equal(got, vo.morph(g, ops)), (g.shape, ops)


def test_morph_golden_closings(dev):
    for name, c in CASES.items():
        if name.endswith("_closed") and name[:-7] in CASES:
            got = dev.morph(dev.BitGrid.from_bool(CASES[name[:-7]]["grid"]), "de").to_bool().cpu().numpy()
            assert sha(got, np.uint8) == c["sha_grid"]


FUSED_CLOSE_SHAPES = [
    (20, 70, 1024), (9, 33, 1000), (3, 29, 600), (1, 5, 160), (2, 1, 33), (17, 300, 128), (40, 9, 31), (5, 2, 64),
    (33, 61, 96), (12, 130, 257), (64, 64, 64), (1, 1, 1), (4, 3, 1024),
]


@pytest.mark.parametrize("shape", FUSED_CLOSE_SHAPES)
@pytest.mark.parametrize("ops", ["de", "dde", "ded", "edde", "dd", "ee", "ed", "eedd"])
def test_fused_closing_tiles(dev, shape, ops):
    """Two consecutive morph operations run as one TMA-staged launch; shapes chosen to cross y tiles, x chunks,
    the periodic wrap inside a tile, partial last words and non-power-of-two row lengths."""
    rng = np.random.default_rng(hash(shape) % 1000)
    for p in (0.02, 0.3, 0.8):
        g = rng.random(shape) < p
        got = dev.morph(dev.BitGrid.from_bool(g), ops).to_bool().cpu().numpy()
        assert np.array_equal(got, vo.morph(g, ops)), (shape, ops, p)


def test_fused_closing_equals_two_steps(dev):
    from mdvcontainment_b200 import synth
    g = synth.salt_noise_grid((192, 200, 224), salt=0.03)
    fused = dev.morph(dev.BitGrid.from_bool(g), "de")
    two = dev.morph(dev.morph(dev.BitGrid.from_bool(g), "d"), "e")          # one launch per operation
    assert torch.equal(fused.words, two.words)
    assert np.array_equal(two.to_bool().cpu().numpy(), vo.morph(g, "de"))


@pytest.mark.parametrize("shape", FUSED_CLOSE_SHAPES)
def test_counted_morph_leaves_the_runs_per_row(dev, shape):
    """mdvc_morph_counted: the TMA-staged kernel (two fused operations, or the odd single one) also counts the z-runs
    of every row of its result (stage 1a of the labelling); mdvc_frame_label_counted on those counts gives the same
    labels as the plain call."""
    import ctypes
    from mdvcontainment_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(7 + sum(shape))
    g = rng.random(shape) < 0.3
    X, Y, Z = shape
    for ops in ("de", "edde", "d", "e", "dde"):
        want = vo.morph(g, ops)
        src = dev.BitGrid.from_bool(g)
        out, tmp = dev.BitGrid(shape), dev.BitGrid(shape)
        plan = dev.FramePlan(shape, max_runs=X * Y * Z + 16, max_labels=X * Y * Z + 16)
        counted = ctypes.c_int(-1)
        rc = lib.mdvc_morph_counted(src.words.data_ptr(), out.words.data_ptr(), tmp.words.data_ptr(), X, Y, Z, src.pitch,
                                    ops.encode(), ctypes.c_void_p(plan.row_counts_ptr()), ctypes.byref(counted),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0 and counted.value == 1            # pairs and the odd single operation both count (rows <= 32 words)
        assert np.array_equal(out.to_bool().cpu().numpy(), want)
        if counted.value:
            runs = 1 + (want[:, :, 1:] != want[:, :, :-1]).sum(axis=2)
            got = plan._view(plan.lib.mdvc_frame_row_counts_ptr, torch.int32, X * Y).cpu().numpy().reshape(X, Y)
            assert np.array_equal(got, runs)
        plan.label(out, rows_counted=bool(counted.value))
        plan.header = plan.read_header()
        assert plan.header.flags == 0
        labels = torch.empty(shape, dtype=torch.int32, device="cuda")
        plan.expand(out, labels, None)
        assert np.array_equal(labels.cpu().numpy(), vo.label_3d_grid(want).astype(np.int32))


def test_morph_rejects_unknown_op(dev):
    from mdvcontainment_b200._lib import MdvcError
    with pytest.raises(MdvcError):
        dev.morph(dev.BitGrid.from_bool(np.zeros((3, 3, 3), bool)), "dx")


# ------------------------------------------------------------------------------------------ A1
@pytest.mark.parametrize("case", ["uniform", "one_plane_range", "two_blobs", "outside_and_far", "on_and_next_to_faces"])
@pytest.mark.parametrize("n", [8192, 50001, 262147])
def test_bin_grid_bucketed_path_equals_plain_binning(dev, case, n):
    """mdvc_bin_grid (bead keys sorted into the key lists of 64 KiB grid slabs, every slab built in shared memory and
    written once; the grid is larger than the 64 MiB below which the call takes the plain kernels) against clear +
    mdvc_bin_atoms and the oracle: same occupancy, voxels and rewritten positions -- for even and very uneven bead
    distributions (a full list spills), bead counts that are no multiple of four, beads far outside the box."""
    rng = np.random.default_rng(n + len(case))
    dims = np.array([5050.0, 3525.0, 4000.0, 90, 90, 90], dtype=np.float32)
    res = 0.5
    _, nbox, T, invT = vo.voxel_transform(dims, res)
    shape = tuple(int(v) for v in np.diagonal(nbox))
    box = dims[:3]
    if case == "uniform":
        pos = rng.random((n, 3)) * box
    elif case == "one_plane_range":
        pos = rng.random((n, 3)) * box * np.array([0.04, 1, 1]) + np.array([0.5 * box[0], 0, 0])
    elif case == "two_blobs":
        pos = np.concatenate([rng.normal(0.2, 0.02, (n // 2, 3)), rng.normal(0.8, 0.05, (n - n // 2, 3))]) * box
    elif case == "on_and_next_to_faces":
        # voxel faces (multiples of the 5 A voxel edge, also outside the box) and their float32 neighbours: the
        # certified single-precision floor must hand every one of them to the float64 path
        faces = (rng.integers(-900, 1800, (n, 3)) * 5.0).astype(np.float32)
        step = rng.integers(-3, 4, (n, 3))
        pos = faces.copy()
        for k in range(1, 4):
            pos = np.where(step >= k, np.nextafter(pos, np.float32(np.inf)), pos)
            pos = np.where(step <= -k, np.nextafter(pos, np.float32(-np.inf)), pos)
        pos[::7] = (rng.random((len(pos[::7]), 3)) * box).astype(np.float32)
        pos[5] = [1e7, -3e6, 2.5]
    else:
        pos = (rng.random((n, 3)) * 7 - 3) * box
    if case != "on_and_next_to_faces":
        pos = np.round(pos, 2)
    pos = pos.astype(np.float32)
    pos_d = torch.from_numpy(pos).cuda()
    plain = dev.BitGrid(shape)
    v0, p0 = dev.bin_atoms(pos_d, T, invT, nbox, plain, want_voxels=True, want_positions=True)
    junk = torch.full((shape[0], shape[1], dev.pitch_words(shape[2])), -1, dtype=torch.int32, device="cuda")
    fresh = dev.BitGrid(shape, words=junk)                      # must be cleared by the call itself
    v1, p1 = dev.bin_grid(pos_d, T, invT, nbox, fresh, want_voxels=True, want_positions=True)
    assert torch.equal(fresh.words, plain.words)
    assert torch.equal(v0, v1) and torch.equal(p0.view(torch.int32), p1.view(torch.int32))
    assert fresh.words.numel() * 4 > 64 << 20                   # the slab path, not the small-grid shortcut
    vox, _, newpos = vo.voxelate_fma(pos, dims, res)
    want_words = np.zeros(tuple(fresh.words.shape), dtype=np.uint32)   # vo.occupancy (atomgroup_to_voxels.py:192-197), bit-packed
    np.bitwise_or.at(want_words, (vox[:, 0], vox[:, 1], vox[:, 2] >> 5), np.uint32(1) << (vox[:, 2] & 31).astype(np.uint32))
    assert np.array_equal(fresh.words.cpu().numpy().view(np.uint32), want_words)
    assert np.array_equal(v1.cpu().numpy(), vox.astype(np.int32))
    # occupancy only, into a reused scratch, unaligned view (falls back to the plain kernels)
    scratch = dev.bin_scratch(n, "cuda")
    again = dev.BitGrid(shape, words=torch.full_like(junk, 7))
    dev.bin_grid(pos_d, T, invT, nbox, again, scratch)
    assert torch.equal(again.words, plain.words)
    odd = dev.BitGrid(shape, words=torch.full_like(junk, 7))
    dev.bin_grid(pos_d[1:], T, invT, nbox, odd, scratch)
    want = dev.BitGrid(shape)
    dev.bin_atoms(pos_d[1:].contiguous(), T, invT, nbox, want)
    assert torch.equal(odd.words, want.words)


def test_bin_grid_beyond_1024_cubed_uses_128k_slabs(dev):
    """A grid of more than 2048 slabs of 64 KiB (1200 x 1200 x 1280 voxels, 220 MiB of bit words) is built from 128 KiB
    slabs, one CTA of the fill kernel per SM: same words as clear + mdvc_bin_atoms."""
    rng = np.random.default_rng(77)
    dims = np.array([6000.0, 6000.0, 6400.0, 90, 90, 90], dtype=np.float32)
    _, nbox, T, invT = vo.voxel_transform(dims, 0.5)
    shape = tuple(int(v) for v in np.diagonal(nbox))
    assert shape == (1200, 1200, 1280)
    pos = (rng.random((150001, 3)) * dims[:3] * 1.2 - 0.1 * dims[:3]).astype(np.float32)
    pos_d = torch.from_numpy(pos).cuda()
    plain = dev.BitGrid(shape)
    dev.bin_atoms(pos_d, T, invT, nbox, plain)
    fresh = dev.BitGrid(shape, words=torch.full((shape[0], shape[1], dev.pitch_words(shape[2])), -1, dtype=torch.int32, device="cuda"))
    dev.bin_grid(pos_d, T, invT, nbox, fresh)
    assert torch.equal(fresh.words, plain.words)
    vox = vo.voxelate_fma(pos, dims, 0.5)[0]
    assert fresh.count() == len(np.unique(vox[:, 0] * (1 << 42) + vox[:, 1] * (1 << 21) + vox[:, 2]))


@pytest.mark.parametrize("name", sorted(binning_cases()))
def test_binning_matches_golden(dev, name):
    c = binning_cases()[name]
    _, nbox, T, invT = vo.voxel_transform(c["dimensions"], c["resolution"])
    pos = torch.from_numpy(c["positions"]).cuda()
    shape = tuple(int(v) for v in np.diagonal(nbox))
    for p in (1, 2, 3):
        bits = dev.BitGrid(shape)
        vox, newpos = dev.bin_atoms(pos, T, invT, nbox, bits, want_voxels=True, want_positions=True)
        assert np.array_equal(vox.cpu().numpy(), c[f"pass{p}_voxels"])
        assert np.array_equal(newpos.cpu().numpy().view(np.uint32), c[f"pass{p}_positions"].view(np.uint32))
        assert np.array_equal(bits.to_bool().cpu().numpy(), c[f"pass{p}_grid"])
        pos = newpos


def test_binning_large_random_vs_oracle(dev):
    rng = np.random.default_rng(8)
    dims = np.array([400.0, 380.0, 255.0, 80, 95, 70], dtype=np.float32)
    n = 300_000
    pos = ((rng.random((n, 3)) * 2.5 - 0.7) * dims[:3]).astype(np.float32)
    pos[::2] = np.round(pos[::2] / 5.0) * 5.0
    vox_o, nbox, new_o = vo.voxelate_fma(pos, dims, 0.5)
    _, _, T, invT = vo.voxel_transform(dims, 0.5)
    shape = tuple(int(v) for v in np.diagonal(nbox))
    bits = dev.BitGrid(shape)
    p = torch.from_numpy(pos).cuda()
    vox, newpos = dev.bin_atoms(p, T, invT, nbox, bits, True, True)
    assert np.array_equal(vox.cpu().numpy().astype(np.int64), vox_o)
    assert np.array_equal(newpos.cpu().numpy().view(np.uint32), new_o.view(np.uint32))
    assert np.array_equal(bits.to_bool().cpu().numpy(), vo.occupancy(vox_o, nbox))
    # in-place write-back (pos_out aliasing pos) gives the same answer
    from mdvcontainment_b200 import _lib
    import ctypes
    lib = _lib.load()
    p2 = torch.from_numpy(pos).cuda()
    Tc, iTc, nb = (np.ascontiguousarray(a) for a in (T, invT, nbox.astype(np.int64)))
    rc = lib.mdvc_bin_atoms(p2.data_ptr(), n, Tc.ctypes.data, iTc.ctypes.data, nb.ctypes.data, None, 0, None,
                            p2.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    assert np.array_equal(p2.cpu().numpy().view(np.uint32), new_o.view(np.uint32))


def test_binning_far_away_atoms_take_the_64_bit_path(dev):
    """Voxel indices beyond 2^28 before wrapping (atoms thousands of box images away) leave the 32-bit fast path."""
    rng = np.random.default_rng(18)
    dims = np.array([60.0, 50.0, 45.0, 90, 90, 90], dtype=np.float32)
    n = 4096
    pos = (rng.random((n, 3)) * dims[:3]).astype(np.float32)
    pos[::3] += np.float32(3.0e10)              # ~6e9 voxels away along every axis
    pos[1::3] -= np.float32(7.5e9)
    vox_o, nbox, new_o = vo.voxelate_fma(pos, dims, 0.5)
    _, _, T, invT = vo.voxel_transform(dims, 0.5)
    shape = tuple(int(v) for v in np.diagonal(nbox))
    bits = dev.BitGrid(shape)
    vox, newpos = dev.bin_atoms(torch.from_numpy(pos).cuda(), T, invT, nbox, bits, True, True)
    assert np.array_equal(vox.cpu().numpy().astype(np.int64), vox_o)
    assert np.array_equal(newpos.cpu().numpy().view(np.uint32), new_o.view(np.uint32))
    assert np.array_equal(bits.to_bool().cpu().numpy(), vo.occupancy(vox_o, nbox))


# ------------------------------------------------------------------------------------------ A5-A8 on runs
def run_frame(dev, g, periodic=True, **plan_kw):
    bits = dev.BitGrid.from_bool(g)
    plan = dev.FramePlan(g.shape, **plan_kw) if plan_kw else None
    plan, h = dev.label_with_retry(bits, plan, periodic=periodic)
    labels = torch.empty(g.shape, dtype=torch.int32, device="cuda")
    plan.expand(bits, labels, None)
    return bits, plan, h, labels


@pytest.mark.parametrize("name", sorted(CASES))
def test_frame_matches_golden(dev, name):
    c = CASES[name]
    g = c["grid"]
    bits, plan, h, labels = run_frame(dev, g)
    assert (h.n_true, h.n_false) == (c["n_pos"], c["n_neg"])
    lab = labels.cpu().numpy()
    assert sha(lab) == c["sha_labels"]
    contacts, bridges = plan.contacts(), plan.bridges()
    assert contacts.shape == c["contacts"].shape and np.array_equal(contacts, c["contacts"])
    assert bridges.shape == c["bridges"].shape and np.array_equal(bridges, c["bridges"])
    vol = plan.label_volumes()
    if c["label_volumes"] is not None:
        for k, v in c["label_volumes"].items():
            assert vol[int(k) + h.n_false] == v
    assert vol.sum() == g.size and vol[h.n_false] == 0
    # components through a host-supplied LUT (the reference's numbering is decided on the host)
    if c["has_arrays"]:
        lut = np.zeros(h.n_true + h.n_false + 1, dtype=np.int32)
        lut[c["labels"].reshape(-1) + h.n_false] = c["components"].reshape(-1)
        plan.set_lut(torch.from_numpy(lut).cuda())
        comps = torch.empty(g.shape, dtype=torch.int32, device="cuda")
        both = torch.empty(g.shape, dtype=torch.int32, device="cuda")
        plan.expand(bits, None, comps)
        assert np.array_equal(comps.cpu().numpy(), c["components"])
        plan.expand(bits, both, comps)          # fused double write
        assert np.array_equal(both.cpu().numpy(), lab) and np.array_equal(comps.cpu().numpy(), c["components"])
    # device-side periodic merge: same partition and counts; numbering by the documented rule
    plan.components(True)
    h2 = plan.read_header()
    lut_dev = plan.lut()
    want_lut, kp, kn = rm.component_lut(bridges, h.n_true, h.n_false)
    assert (h2.n_comp_pos, h2.n_comp_neg) == (kp, kn)
    assert np.array_equal(lut_dev, want_lut)
    cc = plan.component_counts()
    assert cc.sum() == g.size
    assert sorted(cc[cc > 0].tolist()) == sorted(c["counts"].values())


def test_frame_random_shapes_vs_oracle(dev):
    for g in rand_grids(60, 3):
        bits, plan, h, labels = run_frame(dev, g)
        want = vo.label_3d_grid(g).astype(np.int32)
        assert np.array_equal(labels.cpu().numpy(), want), g.shape
        assert np.array_equal(plan.contacts(), vo.find_label_contacts_c(want)), g.shape
        wb = vo.find_bridges_c(want)
        gb = plan.bridges()
        assert gb.shape == wb.shape and np.array_equal(gb, wb), g.shape


def smooth_grid(shape, seed, n_blobs=9, noise=0.0):
    """Union of random periodic ellipsoids and slabs: most neighbouring rows mirror each other (the case the link
    kernels flag as 'nothing for the pair stage to compare'), a few do not, some structures cross the box faces."""
    rng = np.random.default_rng(seed)
    X, Y, Z = shape
    ax = np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij")
    g = np.zeros(shape, bool)
    for _ in range(n_blobs):
        c = rng.random(3) * shape
        r = (0.08 + 0.22 * rng.random(3)) * np.array(shape)
        d = [np.minimum(np.abs(a - ci), n - np.abs(a - ci)) / ri for a, ci, ri, n in zip(ax, c, r, shape)]
        blob = d[0] ** 2 + d[1] ** 2 + d[2] ** 2 < 1.0
        if rng.random() < 0.35:                               # hollow: shells give rows with 5 runs
            blob &= d[0] ** 2 + d[1] ** 2 + d[2] ** 2 > 0.45
        g |= blob
    k = int(rng.choice([i for i in range(3) if shape[i] > 8]))    # one slab through a pair of faces
    sl = [slice(None)] * 3
    lo = int(rng.integers(0, shape[k]))
    sl[k] = slice(lo, lo + max(2, shape[k] // 9))
    g[tuple(sl)] = True
    if noise:
        g ^= rng.random(shape) < noise
    return g


@pytest.mark.parametrize("shape,seed,noise", [((40, 150, 96), 1, 0.0), ((70, 67, 130), 2, 0.0), ((35, 130, 64), 3, 2e-4),
                                              ((1, 200, 90), 4, 0.0), ((66, 1, 70), 5, 0.0), ((34, 129, 33), 6, 1e-3)])
@pytest.mark.parametrize("periodic", [True, False])
def test_frame_smooth_structures_vs_oracle(dev, shape, seed, noise, periodic):
    """Grids of mostly mirrored rows: more than one 64-row strip per plane, more than 32 planes (both link rounds),
    structures through the periodic faces, degenerate extents."""
    g = smooth_grid(shape, seed, noise=noise)
    bits, plan, h, labels = run_frame(dev, g, periodic=periodic)
    want = vo.label_3d_grid(g).astype(np.int32)
    assert np.array_equal(labels.cpu().numpy(), want)
    assert np.array_equal(plan.contacts(), vo.find_label_contacts_c(want))
    if periodic:
        wb = vo.find_bridges_c(want)
        gb = plan.bridges()
        assert gb.shape == wb.shape and np.array_equal(gb, wb)
        # the row list of the pair stage is short on such grids: the flags do skip
        n_listed = int(plan.ws[:256].view(torch.int32)[11])
        assert 0 < n_listed <= shape[0] * shape[1]
        if noise == 0.0 and min(shape[:2]) > 8:
            assert n_listed < 0.6 * shape[0] * shape[1]
    else:
        assert h.n_bridges == 0


def test_frame_slab_mode_skips_bridges(dev):
    g = CASES["complex3D_roll11"]["grid"]
    bits, plan, h, labels = run_frame(dev, g, periodic=False)
    assert h.n_bridges == 0 and plan.bridges().shape == (0,)
    assert np.array_equal(plan.contacts(), CASES["complex3D_roll11"]["contacts"])
    plan.components(False)
    h2 = plan.read_header()
    assert (h2.n_comp_pos, h2.n_comp_neg) == (h.n_true, h.n_false)
    comps = torch.empty(g.shape, dtype=torch.int32, device="cuda")
    plan.expand(bits, None, comps)
    assert np.array_equal(comps.cpu().numpy(), labels.cpu().numpy())


def test_frame_capacity_overflow_is_reported_and_retried(dev):
    g = CASES["salt48"]["grid"]
    bits = dev.BitGrid.from_bool(g)
    tiny = dev.FramePlan(g.shape, max_runs=100, max_labels=8, contact_slots=64, bridge_slots=64)
    tiny.label(bits)
    tiny.pairs(bits)
    h = tiny.read_header()
    assert h.flags & 1 and h.total_runs > 100          # MDVC_OVERFLOW_RUNS with the required size
    plan, h = dev.label_with_retry(bits, tiny)
    assert h.flags == 0
    labels = torch.empty(g.shape, dtype=torch.int32, device="cuda")
    plan.expand(bits, labels, None)
    assert sha(labels.cpu().numpy()) == CASES["salt48"]["sha_labels"]
    assert np.array_equal(plan.contacts(), CASES["salt48"]["contacts"])
    assert np.array_equal(plan.bridges(), CASES["salt48"]["bridges"])


def test_frame_is_deterministic(dev):
    g = CASES["salt_40x24x72"]["grid"]
    out = []
    for _ in range(3):
        bits, plan, h, labels = run_frame(dev, g)
        out.append((sha(labels.cpu().numpy()), sha(plan.contacts()), sha(plan.bridges())))
    assert out[0] == out[1] == out[2]


@pytest.mark.parametrize("n,closing", [(128, True), (256, True)])
def test_hollow_cylinder_vs_oracle(dev, n, closing):
    from mdvcontainment_b200 import synth
    g0 = synth.hollow_cylinder_grid(n)
    bits = dev.morph(dev.BitGrid.from_bool(g0), "de")
    g = bits.to_bool().cpu().numpy()
    assert np.array_equal(g, vo.morph(g0, "de"))
    plan, h = dev.label_with_retry(bits)
    labels = torch.empty(g.shape, dtype=torch.int32, device="cuda")
    plan.expand(bits, labels, None)
    want = vo.label_3d_grid(g).astype(np.int32)
    assert np.array_equal(labels.cpu().numpy(), want)
    assert np.array_equal(plan.contacts(), vo.find_label_contacts_c(want))
    assert np.array_equal(plan.bridges(), vo.find_bridges_c(want))


def test_noise_grid_vs_oracle(dev):
    from mdvcontainment_b200 import synth
    g = synth.salt_noise_grid((96, 80, 160), salt=0.03, seed=9)
    bits, plan, h, labels = run_frame(dev, g)
    want = vo.label_3d_grid(g).astype(np.int32)
    assert np.array_equal(labels.cpu().numpy(), want)
    assert np.array_equal(plan.contacts(), vo.find_label_contacts_c(want))
    assert np.array_equal(plan.bridges(), vo.find_bridges_c(want))
    vals, cnt = np.unique(want, return_counts=True)
    vol = plan.label_volumes()
    assert np.array_equal(vol[vals + h.n_false], cnt)


def test_many_contacts_take_the_multi_block_sort(dev):
    """Pure noise: tens of thousands of labels and more distinct contacts / bridges than one CTA sorts
    (the key capacity grows past 4 * 4096 on retry, which selects the multi-block bitonic network)."""
    rng = np.random.default_rng(33)
    g = rng.random((96, 96, 100)) < 0.03
    bits = dev.BitGrid.from_bool(g)
    plan, h = dev.label_with_retry(bits)
    assert h.flags == 0 and plan.caps.contact_slots > 16384 and h.n_contacts > 16384
    want = vo.label_3d_grid(g).astype(np.int32)
    labels = torch.empty(g.shape, dtype=torch.int32, device="cuda")
    plan.expand(bits, labels, None)
    assert np.array_equal(labels.cpu().numpy(), want)
    assert np.array_equal(plan.contacts(), vo.find_label_contacts_c(want))
    assert np.array_equal(plan.bridges(), vo.find_bridges_c(want))
    # a capacity far above the count: the padded network only touches the live prefix
    big = dev.FramePlan(g.shape, plan.caps.max_runs, plan.caps.max_labels, 1 << 21, 1 << 20)
    big.label(bits)
    big.pairs(bits)
    big.header = big.read_header()
    assert np.array_equal(big.contacts(), plan.contacts()) and np.array_equal(big.bridges(), plan.bridges())


def test_dense_noise_takes_the_global_memory_fallback(dev):
    """More runs per 64-row strip than the shared-memory tiles hold (random voxels, long rows)."""
    rng = np.random.default_rng(21)
    for shape, p in [((6, 70, 256), 0.5), ((3, 130, 320), 0.4), ((40, 3, 700), 0.5)]:
        g = rng.random(shape) < p
        bits, plan, h, labels = run_frame(dev, g)
        want = vo.label_3d_grid(g).astype(np.int32)
        assert np.array_equal(labels.cpu().numpy(), want), shape
        assert np.array_equal(plan.contacts(), vo.find_label_contacts_c(want)), shape
        assert np.array_equal(plan.bridges(), vo.find_bridges_c(want)), shape


# ------------------------------------------------------------------------------------------ seam-level grid functions
@pytest.mark.parametrize("name", [n for n in sorted(CASES) if CASES[n]["has_arrays"]])
def test_grid_contacts_bridges_match_golden(dev, name):
    c = CASES[name]
    got_c = dev.grid_pairs(c["labels"], bridges=False)
    got_b = dev.grid_pairs(c["labels"], bridges=True)
    assert got_c.shape == c["contacts"].shape and np.array_equal(got_c, c["contacts"])
    assert got_b.shape == c["bridges"].shape and np.array_equal(got_b, c["bridges"])


def test_grid_pairs_on_arbitrary_label_grids_with_zeros(dev):
    rng = np.random.default_rng(12)
    for _ in range(20):
        shape = tuple(int(v) for v in rng.integers(1, 9, 3))
        L = rng.integers(-4, 5, shape).astype(np.int32)        # zeros are ignored by the reference
        wc, wb = vo.find_label_contacts_c(L), vo.find_bridges_c(L)
        gc, gb = dev.grid_pairs(L, False), dev.grid_pairs(L, True)
        assert gc.shape == wc.shape and np.array_equal(gc, wc)
        assert gb.shape == wb.shape and np.array_equal(gb, wb)


def test_grid_relabel_count_select(dev):
    rng = np.random.default_rng(13)
    for shape in [(5, 6, 7), (16, 16, 64), (3, 1, 129), (40, 33, 50)]:
        L = rng.integers(-6, 7, shape).astype(np.int32)
        c2l = {3: (1, 2, 5), -1: (-6, -2), 2: (4,), -2: (-1,)}
        want = vo.create_components_grid(L, c2l)
        lut = np.zeros(13, dtype=np.int32)
        for comp, labs in c2l.items():
            for l in labs:
                lut[l + 6] = comp
        got = dev.grid_relabel(L, lut, -6)
        assert np.array_equal(got.cpu().numpy(), want)
        cnt = dev.grid_count(got, -2, 6)
        vals, n = np.unique(want, return_counts=True)
        assert np.array_equal(cnt[vals + 2], n) and cnt.sum() == want.size
        for nodes in ([3], [-1, 2], [0], [3, -2, 2, -1, 0], [9]):
            pos, mask = dev.grid_select(got, nodes, -2, 6, want_mask=True)
            wmask = np.isin(want, nodes)
            assert np.array_equal(mask.cpu().numpy(), wmask)
            assert pos.dtype == np.int64 and np.array_equal(pos, np.array(np.where(wmask)).T.reshape(-1, 3))


# ------------------------------------------------------------------------------------------ atoms
def test_atom_components_and_csr(dev):
    rng = np.random.default_rng(14)
    shape = (20, 17, 40)
    n = 50_000
    vox = np.stack([rng.integers(0, s, n) for s in shape], 1).astype(np.int32)
    comps = rng.integers(-3, 4, shape).astype(np.int32)
    ix = rng.permutation(3 * n)[:n].astype(np.int64)
    vd, cd = torch.from_numpy(vox).cuda(), torch.from_numpy(comps).cuda()
    beta = dev.atom_components(vd, cd).cpu().numpy()
    assert np.array_equal(beta, vo.atom_components(comps, vox))
    csr = dev.AtomCSR(vd, shape)
    lin = (vox[:, 0].astype(np.int64) * shape[1] + vox[:, 1]) * shape[2] + vox[:, 2]
    order = np.argsort(lin, kind="stable")
    assert np.array_equal(csr.order.cpu().numpy(), order.astype(np.int32))
    assert np.array_equal(csr.keys.cpu().numpy(), lin[order])
    mapping = vo.create_mapping(vox, ix)
    for nodes in ([1], [-3, 2], [0, 1, 2, 3, -1, -2, -3], [7]):
        q = vo.voxel_positions(comps, nodes)
        want = vo.voxels2atoms(q, mapping)
        got = csr.select(cd, nodes, -3, 7, torch.from_numpy(ix).cuda())
        assert got.dtype == np.int64 and np.array_equal(got, want)
        # explicit voxel lists: C-order, shuffled, with repeats, empty voxels and out-of-grid entries
        for qq in (q, rng.permutation(q)[:500], np.concatenate([q[:50], q[:50], [[-1, 0, 0], [99, 0, 0]]]) if len(q) else q):
            qq = np.asarray(qq).reshape(-1, 3)
            inside = (qq >= 0).all(1) & (qq < np.array(shape)).all(1)
            assert np.array_equal(csr.gather_voxels(qq, torch.from_numpy(ix).cuda()), vo.voxels2atoms(qq[inside], mapping))


# ------------------------------------------------------------------------------------------ limits and guards
def test_rows_longer_than_the_packed_z_field_are_rejected(dev):
    """Run z-starts are packed into 24 bits in the link / pair kernels: Z >= 2^24 must be refused up front, not
    labelled wrongly (grids that long would pass the 2^31-voxel check, e.g. 4 x 4 x 2^25)."""
    import ctypes
    from mdvcontainment_b200 import _lib
    with pytest.raises(ValueError, match="2\\^24"):
        dev.BitGrid((2, 2, 1 << 24))
    lib = _lib.load()
    Z = 1 << 24
    pitch = lib.mdvc_row_pitch_words(Z)
    words = torch.zeros((1, 1, pitch), dtype=torch.int32, device="cuda")
    caps = _lib.FrameCaps(4096, 4096, 64, 64)
    ws = torch.empty(lib.mdvc_frame_workspace_bytes(1, 1, Z, ctypes.byref(caps)), dtype=torch.uint8, device="cuda")
    rc = lib.mdvc_frame_label(words.data_ptr(), 1, 1, Z, pitch, ws.data_ptr(), ws.numel(), ctypes.byref(caps), None)
    assert rc == _lib.ERR_BADARG
    ok = lib.mdvc_frame_label(words.data_ptr(), 1, 1, Z - 1, pitch, ws.data_ptr(), ws.numel(), ctypes.byref(caps), None)
    assert ok == 0


def test_pipeline_guards():
    """FramePipeline refuses what it cannot do before any device work: non-periodic frames, positions of the wrong
    dtype / layout; FrameOutput.own() detaches the leased grids from the pipeline's buffers."""
    from mdvcontainment_b200 import synth
    from mdvcontainment_b200.pipeline import FramePipeline
    pos, dims = synth.vesicle_scene(box_nm=16.0, seed=2)
    with pytest.raises(NotImplementedError):
        FramePipeline(dims, 0.5, periodic=False)
    pipe = FramePipeline(dims, 0.5, closing=True)
    p = torch.from_numpy(pos).cuda()
    with pytest.raises(ValueError):
        pipe.run(p.double())
    with pytest.raises(ValueError):
        pipe.run(p.t().contiguous().t())
    with pytest.raises(ValueError):
        pipe.run(torch.from_numpy(pos))
    first = pipe.run(p).own()
    keep = first.components_dev.clone()
    other = pipe.run(torch.from_numpy(pos[: len(pos) // 3]).cuda())     # overwrites the pipeline's buffers
    assert torch.equal(first.components_dev, keep)
    assert first.components_dev.data_ptr() != other.components_dev.data_ptr()
    assert not torch.equal(other.components_dev, keep)


def test_staging_memory_is_page_locked_and_outlives_its_views(dev):
    """mdvc_host_alloc through device.staging_empty: pinned (ordinary and write-combined) host tensors usable as H2D
    sources; a slice keeps the allocation alive after the parent tensor is gone."""
    import gc
    for wc in (False, True):
        t = dev.staging_empty((1000, 3), torch.float32, write_combined=wc)
        assert t.shape == (1000, 3) and t.dtype == torch.float32 and t.is_pinned()
        src = torch.arange(3000, dtype=torch.float32).reshape(1000, 3)
        t.copy_(src)
        part = t[10:20]
        del t
        gc.collect()
        d = part.to("cuda", non_blocking=True)
        torch.cuda.synchronize()
        assert torch.equal(d.cpu(), src[10:20])
        del part, d
        gc.collect()


@pytest.mark.parametrize("shape,p", [((5, 7, 1000), 0.5), ((5, 7, 1024), 0.016), ((3, 4, 800), 0.004), ((6, 3, 552), 0.03),
                                     ((2, 9, 1016), 0.9), ((4, 4, 1024), 0.0), ((1, 5, 640), 1.0)])
def test_one_warp_per_row_dense_write(dev, shape, p):
    """Rows of 17..32 words take k_expand_row32 (next row prefetched, run values looked up by warp shuffle when a row has
    at most 32 runs, by loads otherwise): rows with 1, a few, about 32 and hundreds of runs, labels and components,
    each grid alone and both together, against the oracle."""
    from mdvcontainment_b200 import VoxelContainment
    rng = np.random.default_rng(int(1000 * p) + shape[2])
    g = rng.random(shape) < p
    want = vo.label_3d_grid(g).astype(np.int32)
    vc = VoxelContainment(g)
    fr = vc._frame
    both_l = torch.empty(shape, dtype=torch.int32, device="cuda")
    both_c = torch.empty(shape, dtype=torch.int32, device="cuda")
    fr.plan.expand(fr.bits, both_l, both_c)
    only_l = torch.full(shape, 7, dtype=torch.int32, device="cuda")
    fr.plan.expand(fr.bits, only_l, None)
    only_c = torch.full(shape, 7, dtype=torch.int32, device="cuda")
    fr.plan.expand(fr.bits, None, only_c)
    assert np.array_equal(both_l.cpu().numpy(), want)
    assert torch.equal(only_l, both_l) and torch.equal(only_c, both_c)
    lut = torch.as_tensor(fr.lut, device="cuda")
    assert torch.equal(both_c, lut[(both_l + int(fr.header.n_false)).long()])
    runs = 1 + int((g[:, :, 1:] != g[:, :, :-1]).sum(axis=2).max())
    assert (runs > 32) == (p in (0.5, 0.016, 0.03, 0.9))      # both look-up paths are exercised across the cases
