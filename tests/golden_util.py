"""Loaders for tests/golden fixtures (produced by tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a, dtype=None):
    a = np.ascontiguousarray(a if dtype is None else np.asarray(a).astype(dtype))
    return hashlib.sha256(a.tobytes()).hexdigest()[:16]


_vox = None


def voxel_cases():
    """name -> dict(meta..., grid=bool ndarray, contacts, bridges[, labels, components])."""
    global _vox
    if _vox is None:
        with open(os.path.join(GOLD, "voxel_cases.json")) as fh:
            meta = json.load(fh)
        z = np.load(os.path.join(GOLD, "voxel_cases.npz"))
        out = {}
        for name, m in meta.items():
            c = dict(m)
            n = int(np.prod(m["shape"]))
            c["grid"] = np.unpackbits(z[name + "/grid_bits"])[:n].astype(bool).reshape(m["shape"])
            c["contacts"] = z[name + "/contacts"]
            c["bridges"] = z[name + "/bridges"]
            if m["has_arrays"]:
                c["labels"] = z[name + "/labels"]
                c["components"] = z[name + "/components"]
            out[name] = c
        _vox = out
    return _vox


def binning_cases():
    z = np.load(os.path.join(GOLD, "binning_cases.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    out = {}
    for n in names:
        c = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(n + "/")}
        shape = tuple(int(v) for v in c["grid_shape"])
        for p in (1, 2, 3):
            c[f"pass{p}_grid"] = np.unpackbits(c[f"pass{p}_grid_bits"])[: int(np.prod(shape))].astype(bool).reshape(shape)
        c["resolution"] = float(c["resolution"])
        out[n] = c
    return out


def e2e_cases():
    with open(os.path.join(GOLD, "e2e_cases.json")) as fh:
        return json.load(fh)


# hashes computed independently during the survey with the unmodified
# reference.  (grid, labels, contacts, bridges, components)
APPENDIX_D = {
    "complex3D_roll0": ("26ccfeb52da20434", "b6b6889930d6b8ef", "bea875521bfdefa0", "7183aa4df6a15e08", "49dba8c6600881cf"),
    "complex3D_roll3": ("c60898ee91f24520", "f84495cb3c18f9cf", "0583e65063f069fc", "654e47b98d007f62", "4252730b2cceeaa8"),
    "complex3D_roll11": ("a5b47ae25a1d7402", "ae543341a2d20f8b", "0d1acda3cbd8887d", "bdb0eab3470b8cf3", "1f98d79f7133bc85"),
    "complex3D_roll11_closed": ("e0968d4ea59ecd84", "86e0f7c9e5ba6d05", "01d25dc9dd537f09", "c9507d35b0921bbe", "05ba69691628454c"),
    "perlin_32x32x32_s0": ("27ea2387c5832f97", "33b94d165efe338a", "34e74dde90597fa2", "98064bd279b7cc09", "27bd71453981eed5"),
    "perlin_32x32x32_s0_closed": ("e438492c305f7bf0", "34f0aa0e926d1417", "34e74dde90597fa2", "98064bd279b7cc09", "9340c20aca1c726d"),
    "perlin_64x64x64_s1": ("f42cf3a2425b0925", "d67a8fdcf1b71189", "24c60c5cde3744b7", "97c44038830d36aa", "38a9c1431a945872"),
    "perlin_64x64x64_s1_closed": ("8c7692b6c371f5e8", "8945db2267bb2a87", "2f269e777646b012", "802090c7ac483d3d", "688a6cb89e7d3d59"),
    "perlin_48x32x16_s2": ("ecb5eb57c8bc0b06", "c9847cf95192446e", "24c60c5cde3744b7", "0fe84790dbacfeaf", "fd4a16e4eb191c72"),
    "perlin_48x32x16_s2_closed": ("610db60c54d9f813", "d7ba14d0f0a553e4", "18e6876568521546", "9451d7f69b488f3b", "46ca9a30b2efbf35"),
    "perlin_64x64x64_s3": ("d9c9a46b5e4dd7f5", "f2fb41db4bf616f5", "d7e1354389f6c39d", "8d1024a80812ffc8", "4df70f7b642c218b"),
    "perlin_64x64x64_s3_closed": ("53733b6e0502629b", "dc73bb700857a868", "3b3791a372368f57", "afbbc468374ccc01", "843bf302879b27bd"),
}


def tensor_digest(t, chunk=1 << 26):
    """Order-dependent 64-bit checksum (wrapping int64 arithmetic, so the summation order does not matter) of an
    integer / bool torch tensor, computed on the tensor's own device: sum(value[i] * (i mod 1000003 + 1))."""
    import torch
    flat = t.reshape(-1)
    acc = torch.zeros((), dtype=torch.int64, device=flat.device)
    for s in range(0, flat.numel(), chunk):
        part = flat[s:s + chunk].to(torch.int64)
        w = torch.arange(s, s + part.numel(), device=flat.device, dtype=torch.int64) % 1000003 + 1
        acc += (part * w).sum()
    return int(acc.item())


def fullsize_golden(name="fullsize_vesicle_1024.json"):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        return json.load(fh)
