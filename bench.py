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
# reference / CPU baseline leg (the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def _cpu_frame(args):
    """One frame of the reference's CPU algorithm on a scaled copy of the workload. Returns seconds."""
    size, seed = args
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import voxel_oracle as vo
    import refshim
    from mdvcontainment_b200 import host_graph as hg, synth
    pos, dims = synth.vesicle_scene(box_nm=size / 2.0, seed=seed)
    ref_k = refshim.load_ref_kernels()
    t0 = time.perf_counter()
    vox, nbox, _ = vo.voxelate_numpy(pos, dims, 0.5)                     # atomgroup_to_voxels.py:131-142
    grid = vo.occupancy(vox, nbox)                                        # :192-197
    grid = vo.morph(grid, "de")                                           # :278-304
    labels = vo.label_3d_grid(grid).astype(np.int32)                     # label.py:4-19
    uniq = np.unique(labels)                                              # containment_main.py:55
    if ref_k:
        contacts = ref_k["find_label_contacts"].find_label_contacts(labels)   # the reference's own compiled kernels
        bridges = ref_k["find_bridges"].find_bridges(labels)
    else:
        contacts, bridges = vo.find_label_contacts_c(labels), vo.find_bridges_c(labels)
    sol = hg.solve_periodic(uniq, contacts, bridges)
    comps = np.zeros_like(labels)                                         # label.py:32-37, the reference's masked loop
    for comp, labs in sol["component2labels"].items():
        for lab in labs:
            comps[labels == lab] = comp
    counts = dict(zip(*np.unique(comps, return_counts=True)))            # containment_main.py:510-512
    hg.containment_graph(sol["is_contained"], np.unique(comps), sol["component_contact_graph"])
    dt = time.perf_counter() - t0
    return dt, int(grid.size), len(pos), bool(ref_k), len(counts)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = args.cpu_sample
    procs = max(1, min(args.cpu_procs or (os.cpu_count() or 1), 16))
    sys.path[:0] = [os.path.join(ROOT, "oracle")]
    import voxel_oracle as vo
    vo.build_clib()
    import multiprocessing as mp
    pool = mp.get_context("fork").Pool(procs) if procs > 1 else None

    def step(k):
        jobs = [(sample, 7 + k * procs + j) for j in range(procs)]
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
    kind = "reference" if res[0][3] else "port"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 bit-words / int32 labels (fp64 binning)", "data": "synthetic",
        "config": {"workload": f"config 3 vesicle-in-membrane scene scaled to {sample}^3 voxels ({res[0][2]} beads), "
                               f"closing=True; frames/s scaled by voxel count to {args.size}^3",
                   "sample_voxels": n_vox, "processes": procs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind,
                         "sample": f"{procs} independent {sample}^3 frames per step, NumPy/SciPy stages restated from the "
                                   "reference + its compiled Cython contact/bridge kernels (oracle/_ref)" if kind == "reference"
                         else f"{procs} independent {sample}^3 frames per step, oracle port (oracle/_ref not built)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mdvcontainment_b200 import _lib, synth
    from mdvcontainment_b200.pipeline import FramePipeline

    size = args.size
    pos, dims = synth.vesicle_scene(box_nm=size / 2.0, seed=7 + rank)      # every rank its own frame
    n_atoms = len(pos)
    n_flight = max(1, args.in_flight)
    pipes = [FramePipeline(dims, 0.5, closing=True, max_atoms=n_atoms) for _ in range(n_flight)]
    pipe = pipes[0]
    assert pipe.shape == (size, size, size), pipe.shape
    N = size ** 3
    pos_pinned = torch.from_numpy(pos).pin_memory()
    pos_dev = torch.from_numpy(pos).cuda()
    lib = _lib.load()
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = None

    def run_frames(steps, host):
        """`steps` frames with up to n_flight frames in flight: frame i's device stages are enqueued before frame
        i-1's host graph stage runs, so the GPU is never idle waiting for the host."""
        nonlocal out
        pending = []
        for i in range(steps):
            p = pipes[i % n_flight]
            if len(pending) == n_flight:
                out = pending.pop(0).finish()
            if host:
                p.begin(None, host_positions=pos_pinned)
            else:
                p.begin(pos_dev)
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

    # every pipeline in flight gets its warm-up frames (the second one captures its CUDA graph)
    run_frames(max(args.warmup, 3) * n_flight, False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches, expand_ms = timed(args.steps, False)
    clocks = sampler.stop() if rank == 0 else None
    run_frames(max(args.warmup, 3) * n_flight, True)
    ms_e2e, _, _ = timed(args.steps, True)
    # the dominant kernel alone (one frame at a time, nothing else on the GPU), for the roofline's context
    iso = []
    for _ in range(3):
        pipe.run(pos_dev)
        iso.append(pipe.expand_elapsed_ms())
    expand_iso_ms = min(iso)
    # informational: the same frames when only the components grid -- the one dense array the reference's API keeps
    # (containment_main.py:91,134; the non-periodic label grid is a local there) -- is materialised
    for p in pipes:
        p.labels, p.write_labels = None, False
    run_frames(3 * n_flight, False)
    ms_conly, _, _ = timed(args.steps, False)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = world * args.steps / (ms_dev / 1e3)
    e2e_value = world * args.steps / (ms_e2e / 1e3)
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
    small_d2h = 256 + pipe.n_head_c * 8 + pipe.n_head_b * 20 + pipe.n_head_v * 8
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 bit-words / int32 labels (fp64 binning)", "data": "synthetic",
        "config": {"workload": f"config 3: synthetic CG vesicle-in-membrane, {n_atoms} beads, {size}^3 voxels at 0.5 nm, "
                               "closing=True, periodic, labels+components grids written, one frame per step per GPU",
                   "grid": [size] * 3, "atoms": n_atoms, "frames_per_gpu_per_step": 1,
                   "l2": f"no flush: each step streams {frame_bytes / 1e9:.2f} GB (> 126 MB L2)",
                   "frames_in_flight_per_gpu": n_flight,
                   "stage1_cuda_graph": bool(all(len(p._graphs) > 0 for p in pipes)),
                   "parallelism": f"frames sharded over {world} GPU(s), no collective on the data path"},
        "gvoxel_per_s": value * N / 1e9,
        "frame_algorithmic_gbs": frame_bytes * value / world / 1e9,
        "frame_roofline_frac": frame_bytes * value / world / 1e9 / peak,
        "roofline": {"kernel": "k_expand (dense write of labels + components)", "bound": "hbm", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": expand_bytes, "avg_launch_ms": expand_ms,
                     "note": "avg_launch_ms is measured inside the timed region, where the dense write of one frame shares "
                             "the GPU with the run-level kernels of the next frames in flight",
                     "isolated_launch_ms": expand_iso_ms, "isolated_achieved": expand_bytes / (expand_iso_ms / 1e3) / 1e9,
                     "isolated_frac": expand_bytes / (expand_iso_ms / 1e3) / 1e9 / peak},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_atoms * 12),
                "d2h_bytes_per_step": int(small_d2h), "ms_per_step": ms_e2e / args.steps,
                "call": "FramePipeline.begin(host_positions=pinned) / finish() -> containment DAG, ranks, counts on the host"},
        "components_only": {"value": world * args.steps / (ms_conly / 1e3), "unit": UNIT, "ms_per_step": ms_conly / args.steps,
                            "note": "informational, not the headline: labels grid not materialised (4 B/voxel written instead "
                                    "of 8); every other stage identical"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "result": {"components": len(out.voxel_counts), "n_true": out.n_true, "n_false": out.n_false,
                   "dag": str(out).splitlines()[1:8]},
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(args)
        except Exception as exc:            # the GPU line stands on its own
            line["cpu_baseline"] = {"error": repr(exc)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def cpu_baseline(args):
    """Bounded CPU sample of the same workload on this box's host cores (rank 0, N=1 only)."""
    sys.path[:0] = [os.path.join(ROOT, "oracle")]
    import voxel_oracle as vo
    vo.build_clib()
    dt, n_vox, n_pos, have_ref, _ = _cpu_frame((args.cpu_sample, 7))
    value = (1.0 / dt) * n_vox / float(args.size ** 3)
    return {"value": value, "unit": UNIT, "cores": 1, "kind": "reference" if have_ref else "port",
            "seconds_for_sample": dt,
            "sample": f"one {args.cpu_sample}^3 frame of the same scene ({n_pos} beads), frames/s scaled by voxel count to "
                      f"{args.size}^3; NumPy/SciPy stages restated from the reference"
                      + (" + its compiled Cython contact/bridge kernels (oracle/_ref)" if have_ref else " (oracle port)")}


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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
