"""Slab decomposition (mdvcontainment_b200/slab.py) against the single-GPU results / goldens."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from golden_util import sha, voxel_cases

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = voxel_cases()
NAMES = ["complex3D_roll3", "complex3D_roll11", "perlin_64x64x64_s3", "perlin_48x32x16_s2_closed", "rand_33x34x35_p50",
         "rand_12x11x65_p15", "salt48", "cylinder64_closed", "allfalse_4x5x6"]


def run_slabs(grid, n_slabs, morph_ops=""):
    from mdvcontainment_b200 import device as dev
    from mdvcontainment_b200.slab import LocalGather, analyse_slabs, slab_extents
    ext = slab_extents(grid.shape[0], n_slabs)
    bits = {g: dev.BitGrid.from_bool(np.ascontiguousarray(grid[x0:x1])) for g, (x0, x1) in enumerate(ext)}
    res = analyse_slabs(bits, grid.shape, LocalGather(n_slabs), morph_ops=morph_ops)
    labels = np.concatenate([res.labels[g].cpu().numpy() for g in range(n_slabs)], axis=0)
    comps = np.concatenate([res.components[g].cpu().numpy() for g in range(n_slabs)], axis=0)
    return res, labels, comps


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("n_slabs", [1, 2, 3, 4])
def test_slabs_reproduce_the_whole_grid_result(name, n_slabs):
    c = CASES[name]
    g = c["grid"]
    if g.shape[0] < n_slabs:
        pytest.skip("fewer planes than slabs")
    res, labels, comps = run_slabs(g, n_slabs)
    assert sha(labels) == c["sha_labels"]
    assert res.contacts.shape == c["contacts"].shape and np.array_equal(res.contacts, c["contacts"])
    assert res.bridges.shape == c["bridges"].shape and np.array_equal(res.bridges, c["bridges"])
    assert sha(comps) == c["sha_components"]
    assert str(res) == c["str"]
    assert {str(int(k)): int(v) for k, v in res.voxel_counts.items()} == c["counts"]


@pytest.mark.parametrize("n_slabs", [1, 2, 4])
def test_slab_morphology_with_halo_exchange(n_slabs):
    for name in ("perlin_64x64x64_s3", "perlin_48x32x16_s2", "complex3D_roll11"):
        c, closed = CASES[name], CASES[name + "_closed"]
        res, labels, comps = run_slabs(c["grid"], n_slabs, morph_ops="de")
        assert sha(labels) == closed["sha_labels"] and sha(comps) == closed["sha_components"]
        assert str(res) == closed["str"]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


WORKER = textwrap.dedent("""
    import os, sys, json, hashlib
    import numpy as np, torch, torch.distributed as dist
    sys.path[:0] = [%r, %r]
    from golden_util import voxel_cases
    from mdvcontainment_b200 import device as dev
    from mdvcontainment_b200.slab import DistGather, analyse_slabs, slab_extents
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    out = {}
    for name in ("perlin_64x64x64_s3", "salt48", "cylinder64_closed"):
        g = voxel_cases()[name]["grid"]
        x0, x1 = slab_extents(g.shape[0], world)[rank]
        res = analyse_slabs({rank: dev.BitGrid.from_bool(np.ascontiguousarray(g[x0:x1]))}, g.shape, DistGather())
        lab = res.labels[rank].cpu().numpy(); comp = res.components[rank].cpu().numpy()
        parts = [None] * world
        dist.all_gather_object(parts, (lab, comp))
        if rank == 0:
            sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
            out[name] = {"labels": sha(np.concatenate([p[0] for p in parts])), "comps": sha(np.concatenate([p[1] for p in parts])),
                         "str": str(res), "contacts": res.contacts.tolist(), "bridges": res.bridges.tolist()}
    if rank == 0:
        print("RESULT" + json.dumps(out))
    dist.destroy_process_group()
""")


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_slabs(tmp_path):
    import json
    script = tmp_path / "slab_worker.py"
    script.write_text(WORKER % (ROOT, os.path.join(ROOT, "tests")))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    got = json.loads([l for l in out.stdout.splitlines() if l.startswith("RESULT")][0][6:])
    for name, r in got.items():
        c = CASES[name]
        assert r["labels"] == c["sha_labels"] and r["comps"] == c["sha_components"] and r["str"] == c["str"]
        assert r["contacts"] == c["contacts"].tolist() and r["bridges"] == c["bridges"].tolist()


@pytest.mark.parametrize("n_slabs", [2, 4])
def test_generated_slabs_match_single_gpu_containment(n_slabs):
    """The config-5 generator produces the same global grid for any decomposition, and the slab path agrees
    with VoxelContainment on the assembled grid."""
    from mdvcontainment_b200 import VoxelContainment, device as dev, synth
    from mdvcontainment_b200.slab import LocalGather, analyse_slabs, slab_extents
    shape = (96, 80, 128)
    whole = synth.gyroid_salt_planes(0, shape[0], shape, salt=2e-3)
    vc = VoxelContainment(whole)
    ext = slab_extents(shape[0], n_slabs)
    bits = {g: synth.gyroid_salt_planes(x0, x1, shape, salt=2e-3) for g, (x0, x1) in enumerate(ext)}
    assert torch.equal(torch.cat([bits[g].words for g in range(n_slabs)]), whole.words)
    res = analyse_slabs(bits, shape, LocalGather(n_slabs))
    assert str(res) == str(vc)
    assert torch.equal(torch.cat([res.components[g] for g in range(n_slabs)]), vc.components_grid_device)
    assert torch.equal(torch.cat([res.labels[g] for g in range(n_slabs)]), vc.nonp_labeled_grid_device)
    assert res.voxel_counts == vc.voxel_counts
