"""Trajectory driver: frames are independent, so they shard over ranks with no data-path collective
(SURVEY.md 8(e) "Trajectory frames", 8(f) N1).  One process per GPU, launched with torchrun; the only
communication is gathering the tiny per-frame results on rank 0.

    results = run_trajectory(frames, dimensions, resolution)      # every rank calls it

``frames`` is an indexable of float32 [n,3] position arrays (A) -- or a callable ``f -> positions`` plus
``n_frames`` -- and ``results[f]`` is a dict with the frame's DAG text, nodes, edges, ranks and voxel counts.
``dimensions`` is one float32[6] box for the whole trajectory (NVT), a [n_frames, 6] array, or a callable
``f -> float32[6]`` (NPT: the box -- and with it the transform, possibly the grid shape -- changes per frame, as in
the reference, which reads ``atomgroup.dimensions`` in every constructor).
"""
from __future__ import annotations

import os
from typing import Callable, Sequence

import numpy as np


def frames_of_rank(n_frames: int, rank: int, world: int):
    """Frame f is processed by rank f mod world."""
    return list(range(rank, n_frames, world))


class SharedFrameCounter:
    """Frame numbers handed out first come, first served to the ranks of ONE node: an 8-byte counter in a file under
    /dev/shm, incremented under ``flock`` (a few microseconds per claim, no collective, no server).  Ranks whose GPU sits
    behind a slower host link -- on an 8-GPU box the pinned-copy rates differ by 1.5x between ranks
    -- then simply take fewer frames instead of holding everybody up.  ``name`` must be the same on all ranks and new for
    every pass (e.g. derived from MASTER_PORT and a pass number); rank 0 creates the file, the others open it after a
    barrier (``create=False``)."""

    def __init__(self, name: str, create: bool):
        import fcntl
        self._fcntl = fcntl
        self.path = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", f"mdvc_frames_{name}")
        if create:
            fd = os.open(self.path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o600)
            os.pwrite(fd, (0).to_bytes(8, "little"), 0)
        else:
            fd = os.open(self.path, os.O_RDWR)
        self.fd = fd

    def claim(self) -> int:
        """The next unclaimed frame number (monotone over all ranks)."""
        self._fcntl.flock(self.fd, self._fcntl.LOCK_EX)
        try:
            v = int.from_bytes(os.pread(self.fd, 8, 0), "little")
            os.pwrite(self.fd, (v + 1).to_bytes(8, "little"), 0)
        finally:
            self._fcntl.flock(self.fd, self._fcntl.LOCK_UN)
        return v

    def peek(self) -> int:
        return int.from_bytes(os.pread(self.fd, 8, 0), "little")

    def close(self, unlink: bool = False):
        os.close(self.fd)
        if unlink:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def _frame_source(n_frames: int, rank: int, world: int, schedule: str, counter_name: str | None):
    """(next_frame, tail_start, close): ``next_frame()`` -> frame number or None.  ``static``: frame f on rank f mod world.
    ``dynamic``: first come, first served from a SharedFrameCounter (single node); every rank must call this function,
    it synchronises the creation of the counter with a barrier."""
    if schedule == "static" or world == 1:
        it = iter(frames_of_rank(n_frames, rank, world))
        return (lambda: next(it, None)), n_frames, (lambda: None), None
    if schedule != "dynamic":
        raise ValueError(f"schedule must be 'static' or 'dynamic', not {schedule!r}")
    import torch.distributed as dist
    name = counter_name or f"{os.environ.get('MASTER_PORT', '0')}_{_frame_source.passes}"
    _frame_source.passes += 1
    # every rank takes the same path through the collectives whatever happens to its own open(); if any rank cannot
    # reach the counter (ranks on different nodes, no shared /dev/shm) all of them fall back to the static map
    counter, ok = None, True
    if rank == 0:
        try:
            counter = SharedFrameCounter(name, create=True)
        except OSError:
            ok = False
    dist.barrier()
    if rank != 0:
        try:
            counter = SharedFrameCounter(name, create=False)
        except OSError:
            ok = False
    oks = [None] * world
    dist.all_gather_object(oks, ok)
    if not all(oks):
        if counter is not None:
            counter.close(unlink=rank == 0)
        it = iter(frames_of_rank(n_frames, rank, world))
        return (lambda: next(it, None)), n_frames, (lambda: None), None

    def next_frame():
        f = counter.claim()
        return f if f < n_frames else None

    def close():
        dist.barrier()                                   # nobody still claims
        counter.close(unlink=rank == 0)

    return next_frame, n_frames, close, counter


_frame_source.passes = 0


def summarise(out) -> dict:
    """Small, picklable summary of a FrameOutput."""
    return {
        "str": str(out),
        "nodes": out.dag.nodes.tolist(),
        "edges": out.dag.edges(),
        "ranks": {int(k): int(v) for k, v in out.component_ranks.items()},
        "counts": {int(k): int(v) for k, v in zip(out.count_ids.tolist(), out.count_values.tolist())},
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


def dimensions_of_frame(dimensions, f: int) -> np.ndarray:
    """Box of frame ``f`` from a fixed box, a per-frame array or a callable."""
    if callable(dimensions):
        d = dimensions(f)
    else:
        d = np.asarray(dimensions, dtype=np.float32)
        if d.ndim == 2:
            d = d[f]
    d = np.asarray(d, dtype=np.float32).reshape(-1)
    if d.shape != (6,):
        raise ValueError(f"dimensions of frame {f} must be 6 numbers (lengths in A, angles in degrees), got {d.shape}")
    return d


class PipelinePool:
    """``depth`` FramePipelines per grid shape (at most ``max_shapes`` shapes alive: an NPT box that hovers around a
    rounding boundary alternates between two shapes); a box change that keeps the shape only updates the transform."""

    def __init__(self, resolution, closing, morph, depth, max_shapes=2):
        self.args = (resolution, closing, morph)
        self.depth, self.max_shapes = depth, max_shapes
        self.by_shape = {}                                      # shape -> [pipelines]

    def memo_counters(self):
        pipes = [p for ps in self.by_shape.values() for p in ps if p is not None]
        return sum(p.memo_lookups for p in pipes), sum(p.memo_hits for p in pipes)

    def get(self, slot: int, dims: np.ndarray):
        from .pipeline import FramePipeline
        resolution, closing, morph = self.args
        shape = _shape_of(dims, resolution)
        pipes = self.by_shape.get(shape)
        if pipes is None:
            while len(self.by_shape) >= self.max_shapes:        # oldest shape goes (its frames are finished by now)
                self.by_shape.pop(next(iter(self.by_shape)))
            pipes = self.by_shape[shape] = [None] * self.depth
        else:
            self.by_shape[shape] = self.by_shape.pop(shape)      # most recently used last
        if pipes[slot] is None:
            pipes[slot] = FramePipeline(dims, resolution, closing=closing, morph=morph)
        elif not pipes[slot].set_dimensions(dims):
            raise AssertionError("shape bookkeeping out of sync")
        return pipes[slot]


def run_trajectory(frames, dimensions, resolution, closing=True, morph="", n_frames=None,
                   frame_fn: Callable | None = None, in_flight: int = 3, stats: dict | None = None,
                   bind_numa: bool = True, pool: "PipelinePool | None" = None, schedule: str = "static"):
    """Process every frame of a trajectory, sharded over the ranks of the current process group.

    ``schedule="static"``: frame f on rank f mod world.  ``"dynamic"`` (ranks of one node): frames are claimed first come,
    first served from a ``SharedFrameCounter``, so that ranks behind a slower host link -- or frames that cost more --
    do not hold the others up; near the end of the trajectory a rank only claims another frame once it has at most one
    frame left in flight, which leaves the last frames to whoever is ahead.

    Default (CUDA): ``in_flight`` FramePipelines per rank, each with its own streams and a reusable pinned staging
    buffer, so that the copy / device stages of frame f+1.. overlap the host graph stage of frame f.
    A frame that already is a pinned float32 torch tensor is copied to the device straight from where it lies.
    ``frame_fn(positions) -> summary dict`` replaces all of that; tests inject a CPU function to exercise the
    sharding/gathering logic without a GPU.  ``stats`` (a dict) receives this rank's frame count and the hit rate of
    the memoised host graph stage.  ``pool`` (a ``PipelinePool`` made with the same resolution / closing / morph) keeps
    the pipelines -- device buffers, captured CUDA graphs, memoised graph stages -- alive across calls."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if n_frames is None:
        n_frames = len(frames)
    get = frames if callable(frames) else (lambda f: frames[f])
    next_frame, _, close_source, counter = _frame_source(n_frames, rank, world, schedule, None)
    if frame_fn is not None:
        local = {}
        while (f := next_frame()) is not None:
            local[f] = frame_fn(get(f))
        close_source()
        return gather_results(local, n_frames, rank, world)

    import torch
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if bind_numa:                                       # before the pinned staging buffers are allocated
        from .topology import bind_to_gpu
        placement = bind_to_gpu(local_rank)
        if stats is not None:
            stats["placement"] = placement
    if pool is None:
        depth = max(1, min(int(in_flight), -(-n_frames // world) or 1))
        pool = PipelinePool(resolution, closing, morph, depth)
    elif pool.args != (resolution, closing, morph):
        raise ValueError("the pipeline pool was made for another resolution / closing / morph")
    depth = pool.depth
    memo0 = pool.memo_counters()
    pins = [None] * depth
    local, pending = {}, []
    k = -1
    tail_start = n_frames - world * (depth - 1) if counter is not None else n_frames
    while True:
        if counter is not None and len(pending) > 1 and counter.peek() >= tail_start:
            # the last frames go to whoever gets there first with at most one frame left in flight (draining completely
            # would restart the copy / compute pipeline from empty for every one of them)
            f0, p0 = pending.pop(0)
            local[f0] = summarise(p0.finish())
            continue
        f = next_frame()
        if f is None:
            break
        k += 1
        i = k % depth
        if len(pending) == depth:                       # the oldest frame in flight is the one that used slot i
            f0, p0 = pending.pop(0)
            local[f0] = summarise(p0.finish())
        raw = get(f)
        if isinstance(raw, torch.Tensor) and raw.is_pinned() and raw.dtype == torch.float32 and raw.is_contiguous():
            staged = raw.reshape(-1, 3)
        else:
            pos = torch.from_numpy(np.ascontiguousarray(raw, dtype=np.float32).reshape(-1, 3))
            n = pos.shape[0]
            if pins[i] is None or pins[i].shape[0] < n:
                pins[i] = torch.empty((max(n, 1), 3), dtype=torch.float32, pin_memory=True)
            pins[i][:n].copy_(pos)
            staged = pins[i][:n]
        dims = dimensions_of_frame(dimensions, f)
        # a pipeline of a shape that is about to be evicted may still have a frame in flight: drain first
        if _shape_of(dims, resolution) not in pool.by_shape and len(pool.by_shape) >= pool.max_shapes:
            for f0, p0 in pending:
                local[f0] = summarise(p0.finish())
            pending = []
        pipe = pool.get(i, dims)
        pipe.begin(None, host_positions=staged)
        pending.append((f, pipe))
    for f0, p0 in pending:
        local[f0] = summarise(p0.finish())
    close_source()
    if stats is not None:
        lookups, hits = (b - a for a, b in zip(memo0, pool.memo_counters()))
        stats.update(frames=len(local), memo_lookups=lookups, memo_hits=hits, memo_hit_rate=(hits / lookups) if lookups else None)
    return gather_results(local, n_frames, rank, world)


_SHAPES = {}


def _shape_of(dims, resolution):
    """Grid shape of a box (memoised: an NVT trajectory asks for the same box every frame)."""
    key = (np.asarray(dims, dtype=np.float32).tobytes(), float(resolution))
    shape = _SHAPES.get(key)
    if shape is None:
        from .atomgroup_to_voxels import voxel_transform
        _, nbox, _, _ = voxel_transform(dims, resolution)
        if len(_SHAPES) > 4096:
            _SHAPES.clear()
        shape = _SHAPES[key] = tuple(int(v) for v in np.diagonal(nbox))
    return shape
