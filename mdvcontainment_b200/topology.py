"""Host placement of a rank next to its GPU.

The end-to-end path moves 12 B per bead over PCIe every frame (122 MB at 10.2 M beads); with one process per GPU
all ranks pull from host memory at once, so the pinned staging buffers should live on the NUMA node the GPU hangs
off and the copying thread should run there.  ``bind_to_gpu`` does what the process is allowed to do -- CPU affinity
restricted to the node's cores that the container's cpuset permits, and a preferred-node memory policy
(``set_mempolicy``; pinned allocations made afterwards are first-touched locally) -- and reports what happened so
that benchmarks can print it.  Nothing here is needed for correctness.
"""
from __future__ import annotations

import ctypes
import os

_MPOL_PREFERRED = 1
_SYS_set_mempolicy = 238          # x86_64


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int):
    """NUMA node of a CUDA device from sysfs (None when unknown)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu(device_index: int) -> dict:
    """Bind this process (CPU affinity + preferred memory node) to the NUMA node of ``device_index``."""
    info = {"gpu": device_index, "numa_node": gpu_numa_node(device_index), "cpus_bound": None, "mem_policy": None}
    try:
        info["cpus_allowed"] = len(os.sched_getaffinity(0))
    except Exception:
        info["cpus_allowed"] = None
    node = info["numa_node"]
    if node is None:
        return info
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            local = _parse_cpulist(fh.read())
        usable = local & os.sched_getaffinity(0)
        if usable:
            os.sched_setaffinity(0, usable)
            info["cpus_bound"] = len(usable)
        else:
            info["cpus_bound"] = 0                      # the container's cpuset has no core on that node
    except Exception as exc:
        info["cpus_bound"] = f"failed: {exc}"
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        n_bits = 1024
        mask = (ctypes.c_ulong * (n_bits // (8 * ctypes.sizeof(ctypes.c_ulong))))()
        mask[node // (8 * ctypes.sizeof(ctypes.c_ulong))] |= 1 << (node % (8 * ctypes.sizeof(ctypes.c_ulong)))
        rc = libc.syscall(_SYS_set_mempolicy, _MPOL_PREFERRED, ctypes.byref(mask), n_bits)
        info["mem_policy"] = f"preferred node {node}" if rc == 0 else f"failed: errno {ctypes.get_errno()}"
    except Exception as exc:
        info["mem_policy"] = f"failed: {exc}"
    return info
