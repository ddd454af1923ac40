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
                   frame_fn: Callable | None = None):
    """Process every frame of a trajectory, sharded over the ranks of the current process group.

    ``frame_fn(positions) -> summary dict`` defaults to one FramePipeline per rank (CUDA); tests inject
    a CPU function to exercise the sharding/gathering logic without a GPU."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if n_frames is None:
        n_frames = len(frames)
    get = frames if callable(frames) else (lambda f: frames[f])
    if frame_fn is None:
        import torch
        from .pipeline import FramePipeline
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        pipe = FramePipeline(dimensions, resolution, closing=closing, morph=morph)

        def frame_fn(pos):
            t = torch.from_numpy(np.ascontiguousarray(pos, dtype=np.float32)).pin_memory()
            return summarise(pipe.run_host(t))
    local = {f: frame_fn(get(f)) for f in frames_of_rank(n_frames, rank, world)}
    return gather_results(local, n_frames, rank, world)
