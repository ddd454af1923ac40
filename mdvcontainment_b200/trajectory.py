"""Trajectory driver: frames are independent, so they shard over ranks with no data-path collective
(SURVEY.md 8(e) "Trajectory frames", 8(f) N1).  One process per GPU, launched with torchrun; the only
communication is gathering the tiny per-frame results on rank 0.

    results = run_trajectory(frames, dimensions, resolution)      # every rank calls it

``frames`` is an indexable of float32 [n,3] position arrays (A) -- or a callable ``f -> positions`` plus
``n_frames`` -- and ``results[f]`` is a dict with the frame's DAG text, nodes, edges, ranks and voxel counts.
"""
from __future__ import annotations

import os
from typing import Callable, Sequence

import numpy as np


def frames_of_rank(n_frames: int, rank: int, world: int):
    """Frame f is processed by rank f mod world."""
    return list(range(rank, n_frames, world))


def summarise(out) -> dict:
    """Small, picklable summary of a FrameOutput."""
    g = out.containment_graph
    return {
        "str": str(out),
        "nodes": [int(n) for n in g.nodes],
        "edges": [(int(a), int(b)) for a, b in g.edges()],
        "ranks": {int(k): int(v) for k, v in out.component_ranks.items()},
        "counts": {int(k): int(v) for k, v in out.voxel_counts.items()},
    }


def gather_results(local: dict, n_frames: int, rank: int, world: int):
    """All per-frame results on rank 0 (None elsewhere). Uses torch.distributed when world > 1."""
    if world == 1:
        return [local[f] for f in range(n_frames)]
    import torch.distributed as dist
    parts = [None] * world if rank == 0 else None
    dist.gather_object(local, parts, dst=0)
    if rank != 0:
        return None
    merged = {}
    for part in parts:
        merged.update(part)
    return [merged[f] for f in range(n_frames)]


def run_trajectory(frames, dimensions, resolution, closing=True, morph="", n_frames=None,
                   frame_fn: Callable | None = None, in_flight: int = 3):
    """Process every frame of a trajectory, sharded over the ranks of the current process group.

    Default (CUDA): ``in_flight`` FramePipelines per rank, each with its own streams and a reusable pinned staging
    buffer, so that the copy / device stages of frame f+1.. overlap the host graph stage of frame f.
    ``frame_fn(positions) -> summary dict`` replaces all of that; tests inject a CPU function to exercise the
    sharding/gathering logic without a GPU."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if n_frames is None:
        n_frames = len(frames)
    get = frames if callable(frames) else (lambda f: frames[f])
    mine = frames_of_rank(n_frames, rank, world)
    if frame_fn is not None:
        local = {f: frame_fn(get(f)) for f in mine}
        return gather_results(local, n_frames, rank, world)

    import torch
    from .pipeline import FramePipeline
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    depth = max(1, min(int(in_flight), len(mine) or 1))
    pipes = [FramePipeline(dimensions, resolution, closing=closing, morph=morph) for _ in range(depth)]
    pins = [None] * depth
    local, pending = {}, []
    for k, f in enumerate(mine):
        i = k % depth
        if len(pending) == depth:                       # the oldest frame in flight is the one that used slot i
            f0, i0 = pending.pop(0)
            local[f0] = summarise(pipes[i0].finish())
        pos = torch.from_numpy(np.ascontiguousarray(get(f), dtype=np.float32).reshape(-1, 3))
        n = pos.shape[0]
        if pins[i] is None or pins[i].shape[0] < n:
            pins[i] = torch.empty((max(n, 1), 3), dtype=torch.float32, pin_memory=True)
        pins[i][:n].copy_(pos)
        pipes[i].begin(None, host_positions=pins[i][:n])
        pending.append((f, i))
    for f0, i0 in pending:
        local[f0] = summarise(pipes[i0].finish())
    return gather_results(local, n_frames, rank, world)
