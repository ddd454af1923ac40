"""Both frame -> rank maps of run_trajectory on real GPUs (development aid):
torchrun --nproc-per-node 2 tools/traj_dynamic_check.py [frames=24] [box_nm=64]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdvcontainment_b200 import synth  # noqa: E402
from mdvcontainment_b200.trajectory import run_trajectory  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
box = float(sys.argv[2]) if len(sys.argv) > 2 else 64.0
rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
frames = [synth.trajectory_frame(f, box_nm=box) for f in range(n)]
dims = frames[0][1]
res = {}
for schedule in ("static", "dynamic", "dynamic"):
    stats = {}
    res[schedule] = run_trajectory([p for p, _ in frames], dims, 0.5, schedule=schedule, stats=stats, bind_numa=False)
    print(f"rank {rank} {schedule}: {stats.get('frames')} frames", flush=True)
if rank == 0:
    same = all(a["str"] == b["str"] and a["counts"] == b["counts"] for a, b in zip(res["static"], res["dynamic"]))
    print("dynamic == static:", same, "|", res["static"][0]["str"].splitlines()[0])
    assert same and len(res["dynamic"]) == n
if dist.is_initialized():
    dist.destroy_process_group()
