"""Frame sharding of the trajectory driver: world_size-2 gloo run on CPU (no GPU, no data-path collective)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frames_of_rank_partition():
    from mdvcontainment_b200.trajectory import frames_of_rank
    for world in (1, 2, 3, 8):
        seen = sorted(f for r in range(world) for f in frames_of_rank(37, r, world))
        assert seen == list(range(37))
        assert all(f % world == r for r in range(world) for f in frames_of_rank(37, r, world))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    import torch.distributed as dist
    sys.path[:0] = [%r, %r]
    import voxel_oracle as vo
    from mdvcontainment_b200 import host_graph as hg, synth
    from mdvcontainment_b200.trajectory import run_trajectory

    dist.init_process_group("gloo")
    dims = np.array([160, 160, 160, 90, 90, 90], dtype=np.float32)

    def frame(f):
        return synth.trajectory_frame(f, box_nm=16.0)[0]

    def cpu_frame(pos):                      # CPU stand-in for the per-rank FramePipeline (checker code, tests only)
        vox, nbox, _ = vo.voxelate_fma(pos, dims, 0.5)
        grid = vo.morph(vo.occupancy(vox, nbox), "de")
        labels = vo.label_3d_grid(grid).astype(np.int32)
        sol = hg.solve_periodic(np.unique(labels), vo.find_label_contacts_c(labels), vo.find_bridges_c(labels))
        comps = vo.create_components_grid(labels, sol["component2labels"])
        dag = hg.containment_graph(sol["is_contained"], np.unique(comps), sol["component_contact_graph"])
        return {"rank": int(os.environ["RANK"]), "nodes": [int(n) for n in dag.nodes],
                "edges": [(int(a), int(b)) for a, b in dag.edges()],
                "counts": {int(k): int(v) for k, v in vo.counts(comps).items()}}

    res = run_trajectory(frame, dims, 0.5, n_frames=5, frame_fn=cpu_frame, schedule=%r)
    if int(os.environ["RANK"]) == 0:
        print("RESULT" + json.dumps(res))
    dist.destroy_process_group()
""")


def _claim_worker(args):
    from mdvcontainment_b200.trajectory import SharedFrameCounter
    name, n = args
    c = SharedFrameCounter(name, create=False)
    got = [c.claim() for _ in range(n)]
    c.close()
    return got


def test_shared_frame_counter_hands_out_every_number_once():
    """Four processes claiming concurrently: the numbers 0 .. 4n-1, each exactly once, increasing per process."""
    import multiprocessing as mp
    from mdvcontainment_b200.trajectory import SharedFrameCounter
    name = f"test_{os.getpid()}"
    c = SharedFrameCounter(name, create=True)
    try:
        with mp.get_context("fork").Pool(4) as pool:
            parts = pool.map(_claim_worker, [(name, 500)] * 4)
        assert all(p == sorted(p) for p in parts)
        assert sorted(v for p in parts for v in p) == list(range(2000))
        assert c.peek() == 2000 and c.claim() == 2000
    finally:
        c.close(unlink=True)
    assert not os.path.exists(c.path)


@pytest.mark.parametrize("schedule", ["static", "dynamic"])
def test_two_rank_gloo_trajectory_matches_single_rank(tmp_path, schedule):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, os.path.join(ROOT, "oracle"), schedule))

    def launch(world):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert out.returncode == 0, out.stderr[-2000:]
        line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0]
        import json
        return json.loads(line[len("RESULT"):])

    two = launch(2)
    one = launch(1)
    assert len(two) == len(one) == 5
    if schedule == "static":
        assert [r["rank"] for r in two] == [0, 1, 0, 1, 0]        # frame f ran on rank f mod 2
    else:
        assert set(r["rank"] for r in two) <= {0, 1}              # first come, first served
    for a, b in zip(one, two):
        assert a["nodes"] == b["nodes"] and a["edges"] == b["edges"] and a["counts"] == b["counts"]
