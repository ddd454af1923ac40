#!/usr/bin/env python
"""Benchmark of the voxel hot path: frames/s for bin -> close -> periodic CCL -> contacts/bridges ->
components at 1024^3 (BASELINE.json config 3: synthetic CG vesicle-in-membrane, ~10 M beads).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 1024]

One JSON line on stdout (rank 0).  A "step" is one frame.  `value` = frames/s with the bead positions
already resident in HBM; `e2e` = the same through the host-facing pipeline call with pinned host
positions (H2D copy and result read-back inside the timed region).  Frames shard over GPUs with no
data-path collective (weak scaling).  `--impl reference` times the reference's CPU algorithm (oracle
port + the reference's compiled Cython kernels from oracle/_ref when present) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames_per_s_bin_close_ccl_contacts_1024cubed"
UNIT = "frames/s"
B_ALG_PER_VOXEL = 8.5          # bits w 1/8 + closing r/w 1/4 + bits r 1/8 + labels 4 + components 4
B_ALG_PER_ATOM = 12.0
DTYPE = "u32 bit-words / int32 labels (fp64 binning)"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, val in zip(names, r[2:6]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# workload description shared by both arms (the reference arm runs on OUR arm's config)
# ------------------------------------------------------------------------------------------------
SCENES = (                                      # four topologies of the config-3 scene: where the vesicle sits decides
    dict(centre_frac=(0.5, 0.5, 0.625)),        # how many non-periodic labels / bridges it is cut into
    dict(centre_frac=(0.5, 0.5, 0.95)),         # wraps through the z faces
    dict(centre_frac=(0.02, 0.5, 0.625)),       # wraps through the x faces
    dict(centre_frac=(0.98, 0.97, 0.625)),      # wraps through x and y
)


def workload_config(size, n_atoms):
    n = size ** 3
    frame_bytes = B_ALG_PER_VOXEL * n + B_ALG_PER_ATOM * n_atoms
    return {"workload": f"config 3: synthetic CG vesicle-in-membrane, {n_atoms} beads, {size}^3 voxels at 0.5 nm, "
                        "closing=True, periodic, labels+components grids written, one frame per step per GPU; "
                        f"{len(SCENES)} scenes with different topologies rotate through the steps",
            "grid": [size] * 3, "atoms": n_atoms, "frames_per_gpu_per_step": 1,
            "l2": f"no flush: each step streams {frame_bytes / 1e9:.2f} GB (> 126 MB L2)"}


def scene(size, seed, k):
    from mdvcontainment_b200 import synth
    return synth.vesicle_scene(box_nm=size / 2.0, seed=seed, **SCENES[k % len(SCENES)])


# ------------------------------------------------------------------------------------------------
# reference / CPU baseline leg (the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def _load_ref_kernels():
    """The reference's own compiled Cython kernels (oracle/_ref), imported in THIS process -- also in the parent before
    a pool forks, so that whoever watches the parent's loaded shared objects sees them."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import refshim
    return refshim.load_ref_kernels()


def _cpu_frame(args):
    """One frame of the reference's CPU algorithm on a scaled copy of the workload.
    Returns (seconds, voxels, beads, have_ref_kernels, components, per-stage seconds)."""
    size, seed, k = args
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import graph_model as gm
    import voxel_oracle as vo
    pos, dims = scene(size, seed, k)
    ref_k = _load_ref_kernels()
    st = {}
    t0 = t = time.perf_counter()

    def lap(name):
        nonlocal t
        now = time.perf_counter()
        st[name] = st.get(name, 0.0) + now - t
        t = now

    vox, nbox, _ = vo.voxelate_numpy(pos, dims, 0.5)                     # atomgroup_to_voxels.py:131-142
    grid = vo.occupancy(vox, nbox)                                        # :192-197
    lap("create_voxels")
    grid = vo.morph(grid, "de")                                           # :278-304
    lap("morph_voxels")
    labels = vo.label_3d_grid(grid).astype(np.int32)                     # label.py:4-19
    lap("label_3d_grid")
    uniq = np.unique(labels)                                              # containment_main.py:55
    lap("np.unique(labels)")
    if ref_k:
        contacts = ref_k["find_label_contacts"].find_label_contacts(labels)   # the reference's own compiled kernels
        lap("find_label_contacts")
        bridges = ref_k["find_bridges"].find_bridges(labels)
        lap("find_bridges")
    else:
        contacts = vo.find_label_contacts_c(labels)
        lap("find_label_contacts")
        bridges = vo.find_bridges_c(labels)
        lap("find_bridges")
    sol = gm.solve_periodic(uniq, contacts, bridges)                      # graph stage: the oracle's plain-Python model
    c2l = sol["component2labels"]
    lap("graph stage")
    comps = np.zeros_like(labels)                                         # label.py:32-37, the reference's masked loop
    for comp, labs in c2l.items():
        for lab in labs:
            comps[labels == lab] = comp
    lap("create_components_grid")
    counts = dict(zip(*np.unique(comps, return_counts=True)))            # containment_main.py:510-512
    lap("counts")
    gm.containment_graph(sol["is_contained"], np.unique(comps), sol["component_contact_graph"])
    lap("graph stage")
    dt = time.perf_counter() - t0
    return dt, int(grid.size), len(pos), bool(ref_k), len(counts), st


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = args.cpu_sample
    procs = max(1, min(args.cpu_procs or (os.cpu_count() or 1), 16))
    sys.path[:0] = [os.path.join(ROOT, "oracle")]
    import voxel_oracle as vo
    vo.build_clib()
    have_ref = bool(_load_ref_kernels())                                   # in the parent, before forking
    import multiprocessing as mp
    pool = mp.get_context("fork").Pool(procs) if procs > 1 else None

    def step(k):
        jobs = [(sample, 7 + k * procs + j, k * procs + j) for j in range(procs)]
        t0 = time.perf_counter()
        res = pool.map(_cpu_frame, jobs) if pool else [_cpu_frame(jobs[0])]
        return time.perf_counter() - t0, res

    for k in range(args.warmup):
        step(k)
    t_total, res = 0.0, None
    for k in range(args.steps):
        dt, res = step(args.warmup + k)
        t_total += dt
    if pool:
        pool.close()
    n_vox = res[0][1]
    frames_per_s_sample = procs * args.steps / t_total
    value = frames_per_s_sample * n_vox / float(args.size ** 3)          # scaled to the 1024^3 frame of the metric
    kind = "reference" if have_ref and res[0][3] else "port"
    n_atoms_full = int(round(res[0][2] * (args.size / float(sample)) ** 2))   # informational (beads scale ~ area)
    try:
        n_atoms_full = len(scene(args.size, 7, 0)[0])
    except MemoryError:
        pass
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(args.size, n_atoms_full),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind,
                         "sample": f"each step = {procs} independent {sample}^3 frames of the same scenes ({res[0][2]} beads "
                                   f"each) on {procs} processes; frames/s scaled by voxel count to {args.size}^3. NumPy/SciPy "
                                   "stages restated from the reference"
                                   + (" + its compiled Cython contact/bridge kernels (oracle/_ref)" if kind == "reference"
                                      else " (oracle port; oracle/_ref not built)")
                                   + "; graph stage: oracle/graph_model.py (plain Python + networkx)",
                         "sample_voxels": n_vox, "stages_s_last_frame": res[0][5]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def _stats(xs):
    xs = sorted(xs)
    return {"median": statistics.median(xs), "min": xs[0], "max": xs[-1], "n": len(xs)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: this benchmark has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local)
    from mdvcontainment_b200 import topology
    placement = topology.bind_to_gpu(local)               # before any pinned allocation: staging buffers land next to the GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mdvcontainment_b200 import _lib
    from mdvcontainment_b200.pipeline import FramePipeline

    size = args.size
    N = size ** 3
    n_scenes = len(SCENES)
    scenes = [scene(size, 7 + rank, k) for k in range(n_scenes)]           # every rank its own beads
    dims = scenes[0][1]
    n_atoms = len(scenes[0][0])
    assert all(len(p) == n_atoms for p, _ in scenes)
    n_flight = max(1, args.in_flight)
    pipes = [FramePipeline(dims, 0.5, closing=True, max_atoms=n_atoms) for _ in range(n_flight)]
    pipe = pipes[0]
    assert pipe.shape == (size, size, size), pipe.shape
    pos_pinned = [torch.from_numpy(p).pin_memory() for p, _ in scenes]
    pos_dev = [torch.from_numpy(p).cuda() for p, _ in scenes]
    lib = _lib.load()
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = None
    frame_no = [0]

    def run_frames(steps, host):
        """`steps` frames with up to n_flight frames in flight: frame i's device stages are enqueued before frame
        i-1's host graph stage runs, so the GPU is never idle waiting for the host.  Scenes rotate frame by frame."""
        nonlocal out
        pending = []
        for _ in range(steps):
            i = frame_no[0]
            frame_no[0] += 1
            p = pipes[i % n_flight]
            if len(pending) == n_flight:
                out = pending.pop(0).finish()
            if host:
                p.begin(None, host_positions=pos_pinned[i % n_scenes])
            else:
                p.begin(pos_dev[i % n_scenes])
            pending.append(p)
        while pending:
            out = pending.pop(0).finish()

    def timed(steps, host):
        """EXACTLY `steps` frames between barriers; device time by CUDA events, max over ranks."""
        for p in pipes:
            p.expand_elapsed_ms()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.mdvc_launch_count() + sum(p.replayed_launches for p in pipes)
        e0.record()
        run_frames(steps, host)
        for p in pipes:                                   # the default-stream end event must follow all pipelines
            torch.cuda.current_stream().wait_stream(p.stream)
            torch.cuda.current_stream().wait_stream(p.stream_x)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        em = [p.expand_elapsed_ms() for p in pipes]
        em = [v for v in em if v > 0]
        return ms, lib.mdvc_launch_count() + sum(p.replayed_launches for p in pipes) - l0, sum(em) / max(1, len(em))

    claim_pass = [0]
    claimed_frames = [0]

    def timed_claimed(steps):
        """world x steps host-fed frames between barriers, claimed first come, first served from a shared counter
        (trajectory.SharedFrameCounter): ranks behind a slower host link take fewer frames.  Max over ranks."""
        nonlocal out
        from mdvcontainment_b200.trajectory import SharedFrameCounter
        name = f"bench_{os.environ.get('MASTER_PORT', '0')}_{claim_pass[0]}"
        claim_pass[0] += 1
        # every rank takes the same path through the collectives whatever happens to its own open()
        counter, ok = None, 1
        if rank == 0:
            try:
                counter = SharedFrameCounter(name, create=True)
            except OSError:
                ok = 0
        barrier()
        if rank != 0:
            try:
                counter = SharedFrameCounter(name, create=False)
            except OSError:
                ok = 0
        t_ok = torch.tensor([ok], device="cuda")
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        if int(t_ok.item()) == 0:
            if counter is not None:
                counter.close(unlink=rank == 0)
            return None
        total = world * steps
        tail_start = total - world * (n_flight - 1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pending, k = [], 0
        while True:
            if len(pending) > 1 and counter.peek() >= tail_start:   # the last frames go to whoever is ahead
                out = pending.pop(0).finish()
                continue
            i = counter.claim()
            if i >= total:
                break
            p = pipes[k % n_flight]
            k += 1
            if len(pending) == n_flight:
                out = pending.pop(0).finish()
            p.begin(None, host_positions=pos_pinned[i % n_scenes])
            pending.append(p)
        while pending:
            out = pending.pop(0).finish()
        for p in pipes:
            torch.cuda.current_stream().wait_stream(p.stream)
            torch.cuda.current_stream().wait_stream(p.stream_x)
        e1.record()
        barrier()
        counter.close(unlink=rank == 0)
        claimed_frames[0] = k
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def repeated(host, reps):
        """`reps` timed regions of `steps` frames each -> (list of ms, launches of one region, mean dense-write ms)."""
        ms, ln, ex = [], 0, []
        for _ in range(reps):
            m, ln, e = timed(args.steps, host)
            ms.append(m); ex.append(e)
        return ms, ln, sum(ex) / len(ex)

    def set_memo(on):
        for p in pipes:
            p.MEMO_SIZE = 32 if on else 0
            p._solved.clear()
            p._lut_key = None
            p.memo_lookups = p.memo_hits = 0

    reps = max(1, args.repeats)
    # every pipeline in flight sees every scene in its warm-up (the second visit captures the CUDA graph)
    run_frames(max(args.warmup, 3) * n_flight * n_scenes, False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches, expand_ms = repeated(False, reps)
    clocks = sampler.stop() if rank == 0 else None
    run_frames(max(args.warmup, 3) * n_flight, True)
    ms_e2e, _, _ = repeated(True, reps)
    hits = sum(p.memo_hits for p in pipes) / max(1, sum(p.memo_lookups for p in pipes))
    ms_e2e_claimed, claimed_per_rank = None, None
    if world > 1 and timed_claimed(args.steps) is not None:               # warm-up of the claiming loop
        ms_e2e_claimed = [timed_claimed(args.steps) for _ in range(reps)]
        if any(m is None for m in ms_e2e_claimed):
            ms_e2e_claimed = None
        t = torch.zeros(world, device="cuda", dtype=torch.int64)
        t[rank] = claimed_frames[0]
        dist.all_reduce(t)
        claimed_per_rank = [int(v) for v in t.tolist()]
    # the same with the memo of the host graph stage switched off: every frame solves its label graph and uploads its
    # label -> component table (what a trajectory whose topology changes every frame pays)
    set_memo(False)
    run_frames(2 * n_flight, True)
    ms_e2e_miss, _, _ = repeated(True, reps)
    ms_dev_miss, _, _ = repeated(False, max(1, reps // 2))
    set_memo(True)
    run_frames(n_flight * n_scenes, False)

    # plain pinned-host -> device copy ceiling with all ranks copying at once (what bounds the e2e leg), from ordinary
    # page-locked memory and from write-combined page-locked memory
    h2d = h2d_ceiling(torch, dist, world, pos_pinned[0], pipe.pos_dev)
    try:
        from mdvcontainment_b200 import device as dev
        wc = dev.staging_empty(pos_pinned[0].shape, torch.float32, write_combined=True)
        wc.copy_(pos_pinned[0])
        h2d["write_combined"] = h2d_ceiling(torch, dist, world, wc, pipe.pos_dev)
        del wc
    except Exception as exc:                            # noqa: BLE001 - informational measurement
        h2d["write_combined"] = {"error": repr(exc)}

    # the dominant kernel alone (one frame at a time, nothing else on the GPU), for the roofline's context
    iso = []
    for _ in range(3):
        pipe.run(pos_dev[0])
        iso.append(pipe.expand_elapsed_ms())
    expand_iso_ms = min(iso)
    # informational: the same frames when only the components grid -- the one dense array the reference's API keeps
    # (containment_main.py:91,134; the non-periodic label grid is a local there) -- is materialised
    for p in pipes:
        p.labels, p.write_labels = None, False
    run_frames(3 * n_flight, False)
    ms_conly, _, _ = timed(args.steps, False)
    result = {"components": len(out.count_ids), "n_true": out.n_true, "n_false": out.n_false,
              "dag": str(out).splitlines()[1:8]}
    graphs_ok = bool(all(len(p._graphs) > 0 for p in pipes))
    del pipes, pipe, out
    pos_dev = pos_dev[:1]
    torch.cuda.empty_cache()

    extra = {}
    if not args.no_legs:
        extra["trajectory"] = trajectory_leg(torch, dist, rank, world, barrier, h2d_gbs=h2d.get("per_rank_min_gbs"))
        torch.cuda.empty_cache()
        if world > 1:
            extra["slab"] = slab_leg(torch, dist, rank, world, barrier, args.slab_size)
            torch.cuda.empty_cache()
        elif rank == 0:
            extra["api"] = api_leg(torch, scenes[0], size)
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    med = statistics.median(ms_dev)
    med_e2e_static = statistics.median(ms_e2e)
    # N > 1: the end-to-end figure is the better of the two frame -> rank maps (static f mod N, or claimed from a
    # shared counter); both are reported
    med_e2e_claimed = statistics.median(ms_e2e_claimed) if ms_e2e_claimed else None
    use_claimed = med_e2e_claimed is not None and med_e2e_claimed < med_e2e_static
    med_e2e = med_e2e_claimed if use_claimed else med_e2e_static
    if use_claimed:
        ms_e2e = ms_e2e_claimed
    value = world * args.steps / (med / 1e3)
    e2e_value = world * args.steps / (med_e2e / 1e3)
    fps = lambda ms: world * args.steps / (ms / 1e3)                      # noqa: E731
    peak, peak_src = measured_peak()
    traffic, traffic_src = None, None                    # DRAM bytes of one k_expand launch from the committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "expand_traffic.json")) as fh:
            tr = json.load(fh)
        if tr["grid"] == [size] * 3:
            traffic, traffic_src = tr["dram_bytes_read"] + tr["dram_bytes_write"], tr["source"]
    except (OSError, KeyError, ValueError):
        pass
    expand_bytes = N * 8.0 + N / 8.0                      # labels + components written once, bit grid read once
    achieved = expand_bytes / (expand_ms / 1e3) / 1e9
    frame_bytes = B_ALG_PER_VOXEL * N + B_ALG_PER_ATOM * n_atoms
    cfg = workload_config(size, n_atoms)
    small_d2h = 256 + FramePipeline.HEAD_CONTACT_ROWS * 8 + FramePipeline.HEAD_BRIDGE_ROWS * 20 + FramePipeline.HEAD_VOLUMES * 8
    how = {"frames_in_flight_per_gpu": n_flight, "stage1_cuda_graph": graphs_ok,
           "timed_regions": f"{reps} regions of {args.steps} frames each; value = median region (min/max in `spread`)",
           "parallelism": f"frames sharded over {world} GPU(s), no collective on the data path"}
    h2d_bytes = int(n_atoms * 12)
    # every rank moves the same number of frames and the region ends with the slowest rank: the copy bound of this loop is
    # world x the slowest rank's rate; the aggregate rate would need frames to migrate to the ranks with faster links
    e2e_bound = world * h2d["per_rank_min_gbs"] * 1e9 / h2d_bytes if h2d.get("per_rank_min_gbs") else None
    e2e_aggregate_bound = h2d["aggregate_gbs"] * 1e9 / h2d_bytes if h2d.get("aggregate_gbs") else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": med / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic", "config": cfg, "pipeline": how,
        "spread": {"value": {k: (fps(v) if k != "n" else v) for k, v in _stats(ms_dev).items()},
                   "e2e": {k: (fps(v) if k != "n" else v) for k, v in _stats(ms_e2e).items()},
                   "note": "frames/s of the fastest / median / slowest timed region (min <-> max swap: fastest region = max)"},
        "gvoxel_per_s": value * N / 1e9,
        "frame_algorithmic_gbs": frame_bytes * value / world / 1e9,
        "frame_roofline_frac": frame_bytes * value / world / 1e9 / peak,
        "frame_roofline_frac_nominal_8tbs": frame_bytes * value / world / 1e9 / 8000.0,   # north star's nominal HBM3e figure
        # rows of 17..32 words (512 < Z <= 1024) take the one-warp-per-row form of the dense write
        "roofline": {"kernel": ("k_expand_row32" if 16 < (size + 31) // 32 <= 32 and size % 8 == 0 else "k_expand") +
                               " (dense write of labels + components)", "bound": "hbm", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": expand_bytes, "avg_launch_ms": expand_ms,
                     "note": "avg_launch_ms is measured inside the timed region, where the dense write of one frame shares "
                             "the GPU with the run-level kernels of the next frames in flight",
                     "isolated_launch_ms": expand_iso_ms, "isolated_achieved": expand_bytes / (expand_iso_ms / 1e3) / 1e9,
                     "isolated_frac": expand_bytes / (expand_iso_ms / 1e3) / 1e9 / peak},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": int(small_d2h), "ms_per_step": med_e2e / args.steps,
                "call": "FramePipeline.begin(host_positions=pinned) / finish() -> containment DAG, ranks, counts on the host",
                "memo_hit_rate": hits,
                "schedule": ("frames claimed first come, first served from a shared counter (trajectory.SharedFrameCounter)"
                             if use_claimed else "static: every rank processes `steps` frames"),
                "static_sharding_value": world * args.steps / (med_e2e_static / 1e3),
                "claimed_value": (world * args.steps / (med_e2e_claimed / 1e3)) if med_e2e_claimed else None,
                "claimed_frames_per_rank_last_region": claimed_per_rank,
                "memo_miss": {"value": fps(statistics.median(ms_e2e_miss)), "device_resident_value": fps(statistics.median(ms_dev_miss)),
                              "note": "graph-stage memo off: every frame solves its label graph on the host and uploads its "
                                      "label -> component table"},
                "h2d_ceiling_gbs": h2d, "h2d_bound_frames_per_s": e2e_bound,
                "frac_of_h2d_bound": (e2e_value / e2e_bound) if e2e_bound else None,
                "h2d_bound_note": "bound = n_gpus x slowest rank's concurrent pinned-H2D rate / bytes per frame (static sharding: "
                                  "the region ends with the slowest rank); `h2d_aggregate_bound_frames_per_s` is what the sum of "
                                  "all ranks' rates would allow",
                "h2d_aggregate_bound_frames_per_s": e2e_aggregate_bound,
                "frac_of_h2d_aggregate_bound": (e2e_value / e2e_aggregate_bound) if e2e_aggregate_bound else None,
                "placement": placement},
        "components_only": {"value": world * args.steps / (ms_conly / 1e3), "unit": UNIT, "ms_per_step": ms_conly / args.steps,
                            "note": "informational, not the headline: labels grid not materialised (4 B/voxel written instead "
                                    "of 8); every other stage identical"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "result": result,
    }
    line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(args)
        except Exception as exc:            # the GPU line stands on its own
            line["cpu_baseline"] = {"error": repr(exc)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def h2d_ceiling(torch, dist, world, pinned, dev_buf, n_copies=20):
    """GB/s of a plain pinned-host -> device copy of one frame's positions, all ranks copying concurrently."""
    n = pinned.shape[0]
    dst = dev_buf[:n] if dev_buf is not None and dev_buf.shape[0] >= n else torch.empty_like(pinned, device="cuda")
    for _ in range(3):
        dst.copy_(pinned, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_copies):
        dst.copy_(pinned, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = n_copies * pinned.numel() * 4 / (e0.elapsed_time(e1) / 1e3) / 1e9
    if world > 1:
        t = torch.tensor([gbs, gbs, gbs], device="cuda", dtype=torch.float64)
        lo, hi, tot = t[0:1].clone(), t[1:2].clone(), t[2:3].clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        return {"per_rank_min_gbs": float(lo.item()), "per_rank_max_gbs": float(hi.item()), "aggregate_gbs": float(tot.item()),
                "ranks_concurrent": world, "bytes_per_copy": int(pinned.numel() * 4)}
    return {"per_rank_min_gbs": gbs, "per_rank_max_gbs": gbs, "aggregate_gbs": gbs, "ranks_concurrent": 1,
            "bytes_per_copy": int(pinned.numel() * 4)}


def trajectory_leg(torch, dist, rank, world, barrier, frames_per_rank=64, box_nm=256.0, h2d_gbs=None):
    """Config 4: distinct frames of the synthetic 512^3 trajectory (vesicle drifting through the periodic box and
    breathing, fresh bead jitter per frame), sharded f -> rank f mod world, through trajectory.run_trajectory."""
    from mdvcontainment_b200 import synth
    from mdvcontainment_b200.trajectory import PipelinePool, run_trajectory
    n_frames = frames_per_rank * world
    mine = list(range(rank, n_frames, world))
    dims = None
    pinned = {}
    for f in mine:
        pos, dims = synth.trajectory_frame(f, box_nm=box_nm)
        pinned[f] = torch.from_numpy(pos).pin_memory()
    beads = sum(int(p.shape[0]) for p in pinned.values()) / max(1, len(pinned))

    pool = PipelinePool(0.5, True, "", 3)                      # one trajectory session: buffers and CUDA graphs are kept

    def run(stats=None):
        return run_trajectory(lambda f: pinned[f], dims, 0.5, closing=True, n_frames=n_frames, stats=stats, bind_numa=False,
                              pool=pool)

    def forget():
        """Every timed pass starts like a fresh trajectory: nothing memoised from the previous pass."""
        for ps in pool.by_shape.values():
            for p in ps:
                if p is not None:
                    p._solved.clear()
                    p._lut_key = None

    run()                                                     # warm-up: allocations, CUDA graph capture
    times, stats, res = [], {}, None
    for _ in range(3):
        stats = {}
        forget()
        barrier()
        t0 = time.perf_counter()
        res = run(stats)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        times.append(dt)
    out = {"workload": f"config 4: {n_frames} distinct frames of synth.trajectory_frame at {int(2 * box_nm)}^3 voxels "
                       f"(~{int(beads)} beads each), pinned host positions, results gathered on rank 0",
           "frames": n_frames, "frames_per_s": n_frames / statistics.median(times),
           "frames_per_s_min_max": [n_frames / max(times), n_frames / min(times)],
           "memo_hit_rate": stats.get("memo_hit_rate"),
           "memo_note": "graph-stage memo emptied before every timed pass: hits are repeats of a topology within the pass",
           "timing": "wall clock around run_trajectory incl. the result gather, max over ranks"}
    if h2d_gbs:
        # the positions cross PCIe once per frame: what the slowest rank's measured pinned-copy rate allows
        out["h2d_bytes_per_frame"] = int(beads * 12)
        out["h2d_bound_frames_per_s"] = world * h2d_gbs * 1e9 / (beads * 12)
        out["frac_of_h2d_bound"] = out["frames_per_s"] / out["h2d_bound_frames_per_s"]
    if rank == 0 and res is not None:
        out["distinct_component_counts"] = sorted({len(r["nodes"]) for r in res})
    return out


def slab_leg(torch, dist, rank, world, barrier, n):
    """Config 5's mode on the data path that has a collective: ONE n^3 many-label grid cut into `world` x-slabs, one
    per rank, analysed through slab.analyse_slabs with NCCL gathers (boundary planes, pair lists); checked against
    the whole grid analysed on rank 0 alone by digests of the label and component grids."""
    from mdvcontainment_b200 import device as dev, synth
    from mdvcontainment_b200.pipeline import FramePipeline
    from mdvcontainment_b200.slab import DistGather, analyse_slabs, slab_extents
    shape = (n, n, n)
    salt = 1e-3
    x0, x1 = slab_extents(n, world)[rank]
    bits = synth.gyroid_salt_planes(x0, x1, shape, salt=salt)
    res = analyse_slabs({rank: bits}, shape, DistGather())                  # warm-up: the workspaces find their capacities
    plans = res.plans
    times = []
    for _ in range(3):
        del res
        barrier()
        t0 = time.perf_counter()
        res = analyse_slabs({rank: bits}, shape, DistGather(), plans=plans)
        barrier()
        times.append(time.perf_counter() - t0)
    dt = statistics.median(times)
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    off = x0 * n * n
    dig = torch.tensor([dev.tensor_digest(res.labels[rank], off), dev.tensor_digest(res.components[rank], off)],
                       device="cuda", dtype=torch.int64)
    dist.all_reduce(dig, op=dist.ReduceOp.SUM)                             # wrapping int64: partial digests add up
    out = {"workload": f"config 5 mode: one {n}^3 grid, periodic gyroid XOR Bernoulli({salt}) noise, closing=False, "
                       f"{world} x-slabs over NCCL", "grid": [n] * 3, "ranks": world, "seconds": dt,
           "gvoxel_per_s": n ** 3 / dt / 1e9, "labels": int(res.n_true + res.n_false), "components": int(len(res.dag)),
           "contacts": int(len(res.contacts)), "bridges": int(len(np.asarray(res.bridges).reshape(-1, 5))),
           "counts_sum_ok": bool(int(res.count_values.sum()) == n ** 3), "stages_s": {k: round(v, 4) for k, v in res.timings.items()},
           "collective": "NCCL all_gather (planes) + all_gather_object (pairs)"}
    n_contacts, n_bridges = out["contacts"], out["bridges"]
    del res
    torch.cuda.empty_cache()
    if rank == 0 and n ** 3 < 2 ** 31:
        whole = synth.gyroid_salt_planes(0, n, shape, salt=salt)
        dims = np.array([n * 5.0] * 3 + [90, 90, 90], dtype=np.float32)
        single = FramePipeline(dims, 0.5, closing=False, use_graph=False).analyse(whole)
        torch.cuda.synchronize()
        want = [dev.tensor_digest(single.labels_dev), dev.tensor_digest(single.components_dev)]
        out["parity_ok"] = bool(want == [int(v) for v in dig.tolist()] and len(single.contacts) == n_contacts
                                and len(np.asarray(single.bridges).reshape(-1, 5)) == n_bridges)
        out["parity"] = "label and component grids: 64-bit digests of all slabs == single-GPU whole grid (rank 0)"
    barrier()
    return out


def api_leg(torch, scene0, size):
    """The drop-in constructors a reference user calls, wall clock at full size (one object at a time)."""
    from mdvcontainment_b200 import Containment
    from mdvcontainment_b200.universe import Universe
    pos, dims = scene0
    out = {"grid": [size] * 3, "atoms": len(pos), "timing": "wall clock, median of 3 after one warm-up call"}
    import contextlib
    import io
    for name, kw in (("no_mapping", dict(no_mapping=True, betafactors=False)), ("default_flags", dict())):
        ts = []
        for k in range(4):
            u = Universe(pos.copy(), dims)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                c = Containment(u.atoms, 0.5, closing=True, **kw)
                text = str(c)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
            del c, u
        out[name + "_ms"] = 1e3 * statistics.median(ts[1:])
        out[name + "_first_line"] = text.splitlines()[1] if len(text.splitlines()) > 1 else text
    return out


def cpu_baseline(args):
    """Bounded CPU sample of the same workload on this box's host cores (rank 0, N=1 only): ONE 512^3 frame on one
    core with the per-stage split of BASELINE.md section 4."""
    sys.path[:0] = [os.path.join(ROOT, "oracle")]
    import voxel_oracle as vo
    vo.build_clib()
    sample = args.cpu_baseline_size
    dt, n_vox, n_pos, have_ref, _, stages = _cpu_frame((sample, 7, 0))
    value = (1.0 / dt) * n_vox / float(args.size ** 3)
    return {"value": value, "unit": UNIT, "cores": 1, "kind": "reference" if have_ref else "port",
            "seconds_for_sample": dt, "measured": {"grid": [sample] * 3, "beads": n_pos, "frames_per_s_at_that_size": 1.0 / dt},
            "stages_s": {k: round(v, 4) for k, v in stages.items()},
            "sample": f"one {sample}^3 frame of the same scene ({n_pos} beads) on one core; `value` = its frames/s scaled by "
                      f"voxel count to {args.size}^3 (`measured` is the unscaled figure); NumPy/SciPy stages restated from the "
                      "reference" + (" + its compiled Cython contact/bridge kernels (oracle/_ref)" if have_ref else " (oracle port)")
                      + "; graph stage: oracle/graph_model.py"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--cpu-sample", type=int, default=256, help="grid side of the bounded CPU sample")
    ap.add_argument("--cpu-procs", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--in-flight", type=int, default=3, help="frames in flight per GPU (pipelines on separate streams)")
    ap.add_argument("--repeats", type=int, default=5, help="timed regions of --steps frames each (median reported)")
    ap.add_argument("--cpu-baseline-size", type=int, default=512, help="grid side of the one-core CPU baseline frame")
    ap.add_argument("--slab-size", type=int, default=1024, help="grid side of the slab-decomposition leg (N > 1)")
    ap.add_argument("--no-legs", action="store_true", help="skip the trajectory / slab / api legs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
