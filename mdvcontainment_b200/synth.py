"""Deterministic synthetic inputs for BASELINE.json's configs.

Host-side NumPy only (inputs are *data*, not part of the hot path).  Positions are in
Angstrom, float32, rounded to 0.01 A (the precision of a ``.gro`` file) so that a realistic
share of beads sits exactly on voxel faces.
"""
from __future__ import annotations

import numpy as np

BEAD_DENSITY = 6.25          # Martini tail beads per nm^3


def hollow_cylinder_grid(n: int, nz: int | None = None) -> np.ndarray:
    """Config 2: hollow cylinder periodic along z, the scaled form of the cylinder in the
    reference's ``gen_data.py:13-14`` -- True where 0.2 n <= r < 0.3 n."""
    nz = n if nz is None else nz
    c = np.arange(n) - n / 2 + 0.5
    r = np.hypot(c[:, None], c[None, :])
    ring = (r < 0.3 * n) & (r >= 0.2 * n)
    return np.broadcast_to(ring[:, :, None], (n, n, nz)).copy()


def _shell(rng, centre_nm, radius_nm, thickness_nm, density):
    lo, hi = radius_nm - thickness_nm / 2, radius_nm + thickness_nm / 2
    n = int(round(4.0 / 3.0 * np.pi * (hi ** 3 - lo ** 3) * density))
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.cbrt(rng.uniform(lo ** 3, hi ** 3, n))
    return np.asarray(centre_nm) + d * r[:, None]


def vesicle_scene(box_nm: float = 512.0, seed: int = 7, density: float = BEAD_DENSITY,
                  centre_frac=(0.5, 0.5, 0.625), radius_frac: float = 100.0 / 512.0,
                  inner_frac: float = 40.0 / 512.0, thickness_nm: float = 4.0,
                  slab_z_frac: float = 0.25):
    """Config 3: planar bilayer slab + vesicle + nested vesicle in a cubic box.

    box 512 nm at 0.5 nm resolution is the 1024^3 case (about 10.2 M beads); smaller boxes
    scale the radii with the box and keep the 4 nm leaflet thickness.
    Returns (positions float32 [n,3] in A, dimensions float32 [6])."""
    rng = np.random.default_rng(seed)
    n_slab = int(round(box_nm * box_nm * thickness_nm * density))
    slab = rng.random((n_slab, 3)) * np.array([box_nm, box_nm, thickness_nm])
    slab[:, 2] += slab_z_frac * box_nm - thickness_nm / 2
    centre = np.asarray(centre_frac) * box_nm
    outer = _shell(rng, centre, radius_frac * box_nm, thickness_nm, density)
    inner = _shell(rng, centre, inner_frac * box_nm, thickness_nm, density)
    pos_nm = np.concatenate([slab, outer, inner])
    pos = np.round(pos_nm * 10.0, 2).astype(np.float32)
    dims = np.array([box_nm * 10] * 3 + [90, 90, 90], dtype=np.float32)
    return pos, dims


def trajectory_frame(f: int, box_nm: float = 256.0, density: float = BEAD_DENSITY):
    """Config 4: frame ``f`` of the synthetic trajectory -- the config-3 scene at half scale,
    vesicle centre drifting through the periodic box and radius breathing."""
    drift = np.array([0.37, 0.11, 0.23]) * f / box_nm
    centre = (np.array([0.5, 0.5, 0.625]) + drift) % 1.0
    breathe = 1.0 + 0.05 * np.sin(2 * np.pi * f / 100.0)
    return vesicle_scene(box_nm, seed=1000 + f, density=density, centre_frac=tuple(centre),
                         radius_frac=100.0 / 512.0 * breathe)


def bicelle_scene(seed: int = 2024):
    """Config 1(ii): CG-like bicelle disc (selection) in solvent.

    Returns (positions A float32, dimensions float32[6], names) with names 'C' for the
    disc beads and 'W' for solvent."""
    rng = np.random.default_rng(seed)
    box = np.array([30.0, 30.0, 20.0])
    n_cand = int(np.pi * 8.0 ** 2 * 4.0 * BEAD_DENSITY * 4 / np.pi) + 1      # bounding square prism
    cand = rng.random((n_cand, 3)) * np.array([16.0, 16.0, 4.0]) + np.array([7.0, 7.0, 8.0])
    disc = cand[np.hypot(cand[:, 0] - 15.0, cand[:, 1] - 15.0) < 8.0]
    n_w = int(box.prod() * 8.3)
    w = rng.random((n_w, 3)) * box
    inside = (np.hypot(w[:, 0] - 15.0, w[:, 1] - 15.0) < 8.0) & (np.abs(w[:, 2] - 10.0) < 2.0)
    w = w[~inside]
    pos = np.round(np.concatenate([disc, w]) * 10.0, 2).astype(np.float32)
    names = np.array(["C"] * len(disc) + ["W"] * len(w), dtype=object)
    dims = np.array([300, 300, 200, 90, 90, 90], dtype=np.float32)
    return pos, dims, names


def salt_noise_grid(shape, p_true: float = 0.5, salt: float = 0.02, seed: int = 5, blob: int = 16):
    """Config 5 stand-in (many labels): periodic low-frequency blobs XOR Bernoulli(salt) noise."""
    rng = np.random.default_rng(seed)
    coarse = rng.random(tuple(max(1, s // blob) for s in shape))
    for ax, s in enumerate(shape):
        reps = -(-s // coarse.shape[ax])
        coarse = np.repeat(coarse, reps, axis=ax).take(np.arange(s), axis=ax)
    for ax in range(3):   # periodic smoothing
        coarse = (coarse + np.roll(coarse, blob // 2, ax) + np.roll(coarse, -(blob // 2), ax)) / 3
    base = coarse > np.quantile(coarse, 1 - p_true)
    return base ^ (rng.random(shape) < salt)


def gyroid_salt_planes(x0: int, x1: int, shape, salt: float = 1e-4, seed: int = 5, device="cuda", chunk: int = 8):
    """Config 5 at scale: planes [x0, x1) of a periodic gyroid-like solid (sin-sum level set, periodic in all
    three axes) XOR Bernoulli(salt) noise, generated on the device plane-chunk by plane-chunk and returned as a
    bit-packed ``BitGrid`` of the slab.  Every plane depends only on its global x and the seed, so any slab
    decomposition produces the same global grid (input generation, not part of the hot path)."""
    import torch
    from . import device as dev
    X, Y, Z = (int(v) for v in shape)
    out = dev.BitGrid((x1 - x0, Y, Z), device=device)
    yy = torch.arange(Y, device=device, dtype=torch.float32)[None, :, None] * (2 * np.pi * 3 / Y)
    zz = torch.arange(Z, device=device, dtype=torch.float32)[None, None, :] * (2 * np.pi * 2 / Z)
    for c0 in range(x0, x1, chunk):
        c1 = min(c0 + chunk, x1)
        xx = torch.arange(c0, c1, device=device, dtype=torch.float32)[:, None, None] * (2 * np.pi * 2 / X)
        solid = (torch.sin(xx) * torch.cos(yy) + torch.sin(yy) * torch.cos(zz) + torch.sin(zz) * torch.cos(xx)) > 0.3
        if salt > 0:
            noise = torch.empty((c1 - c0, Y, Z), device=device, dtype=torch.float32)
            for k, x in enumerate(range(c0, c1)):
                gen = torch.Generator(device=device)
                gen.manual_seed(seed * 1_000_003 + x)
                noise[k].uniform_(generator=gen)
            solid ^= noise < salt
        part = dev.BitGrid.from_bool(solid)
        out.words[c0 - x0:c1 - x0].copy_(part.words)
    return out
