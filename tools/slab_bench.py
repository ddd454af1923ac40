"""Slab-decomposed containment of one oversized periodic grid (BASELINE.json config 5).

    torchrun --nnodes=1 --nproc-per-node G tools/slab_bench.py [--size 2048] [--salt 1e-4]

Each rank generates its x-slab of a synthetic periodic solid + salt noise on its GPU, then all ranks run
mdvcontainment_b200.slab.analyse_slabs (NCCL all-gather of boundary planes over NVLink).  Prints one JSON line.
"""
import argparse, json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdvcontainment_b200 import synth
from mdvcontainment_b200.slab import DistGather, LocalGather, analyse_slabs, slab_extents

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=2048)
ap.add_argument("--x", type=int, default=0, help="number of planes (default: size)")
ap.add_argument("--salt", type=float, default=1e-3)
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
shape = (args.x or args.size, args.size, args.size)
x0, x1 = slab_extents(shape[0], world)[rank]
t0 = time.perf_counter()
bits = synth.gyroid_salt_planes(x0, x1, shape, salt=args.salt)
torch.cuda.synchronize()
t_gen = time.perf_counter() - t0
gather = DistGather() if world > 1 else LocalGather(1)
times = []
for _ in range(args.reps):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = analyse_slabs({rank: bits}, shape, gather)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times.append(time.perf_counter() - t0)
    labels, comps = res.labels[rank], res.components[rank]
n_vox = float(shape[0]) * shape[1] * shape[2]
total_counts = int(res.count_values.sum())
local_hist_ok = int(torch.count_nonzero(comps == 0).item()) == 0
if rank == 0:
    print(json.dumps({
        "workload": f"config 5: periodic solid + Bernoulli({args.salt}) salt, {shape[0]}x{shape[1]}x{shape[2]} voxels, "
                    f"{world} x-slab(s), closing=False", "n_gpus": world, "voxels": n_vox,
        "beyond_reference_limit": bool(n_vox >= 2 ** 31 - 2),
        "seconds_per_grid": min(times), "gvoxel_per_s": n_vox / min(times) / 1e9, "generate_s": t_gen,
        "labels": [res.n_true, res.n_false], "contacts": int(len(res.contacts)), "bridges": int(len(np.asarray(res.bridges).reshape(-1, 5))),
        "components": int(len(res.dag)), "dag_edges": int(len(res.dag.edge_src)), "roots": len(res.dag.roots()),
        "largest_components": sorted(((int(v), int(k)) for k, v in zip(res.count_ids[np.argsort(res.count_values)[-4:]],
                                                                       np.sort(res.count_values)[-4:])), reverse=True),
        "counts_sum_equals_voxels": total_counts == int(n_vox), "no_unlabelled_voxels_rank0": local_hist_ok,
        "all_times_s": times, "timings_s": getattr(res, "timings", None)}))
if world > 1:
    dist.destroy_process_group()
