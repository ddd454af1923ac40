"""Store-pattern micro-benchmark for the dense write (development aid): python tools/microbench/write_pattern.py"""
import ctypes
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "write_pattern.so")
if not os.path.exists(SO):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "-Xcompiler", "-fPIC",
                           os.path.join(HERE, "write_pattern.cu"), "-o", SO])
lib = ctypes.CDLL(SO)
lib.write_pattern.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                              ctypes.c_void_p]
n = 1024
a = torch.empty((n, n, n), dtype=torch.int32, device="cuda")
b = torch.empty((n, n, n), dtype=torch.int32, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
lib.write_pattern_slabs.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
for n_slabs in (4,):
    def run_slabs():
        assert lib.write_pattern_slabs(a.data_ptr(), b.data_ptr(), n * n, n, n_slabs, st) == 0
    for _ in range(2):
        run_slabs()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run_slabs()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"two grids + a 128 B read per row, input prefetched into the L2 slab by slab, {n_slabs:3d} slabs: {ms * 1e3:8.1f} us  "
          f"{n ** 3 * 8 / ms / 1e6:7.0f} GB/s")
cases = [(mode, 256, 0) for mode in (0, 4, 16, 18, 19, 15, 17, 13, 14, 8, 9, 10, 11, 12, 5, 6)] + [(mode, ctas, 0) for mode in (0, 1, 2, 3) for ctas in (256, 16, 8, 4)]
# non-persistent grid, RESIDENT CTAs per SM limited through dynamic shared memory (227 KB per SM): 8, 6, 5, 4, 3, 2
cases += [(0, 256, kb * 1024) for kb in (24, 36, 44, 56, 72, 110)] + [(2, 256, kb * 1024) for kb in (24, 44, 56, 110)]
names = {18: "512 B read (one 16-byte load per lane) every 4th row", 19: "1 KiB read (two 16-byte loads per lane) every 8th row",
         15: "32 B read per row", 16: "128 B read every 4th row", 17: "128 B read every 16th row",
         13: "read per row from the same 4 KiB (no DRAM read)", 14: "read per row from the same 2 MiB, ld.cg (L2 hits, no DRAM read)",
         8: "read per row NOT feeding the stores", 9: "read per row with ld.global.cg", 10: "read per row with ld.global.nc",
         11: "next row's word loaded while this row is stored", 12: "volatile read per row",
         4: "two grids + a 128 B read per row", 5: "two grids + ~60 dependent instructions per store", 6: "two grids + read + instructions",
         0: "two grids, interleaved per quarter row", 1: "two grids, row by row", 2: "one grid", 3: "two grids, blocked rows per warp"}
for mode, ctas, smem in cases:
    name = names[mode] + (f", {smem // 1024} KB smem per CTA" if smem else "")
    if True:
        def run():
            assert lib.write_pattern(a.data_ptr(), b.data_ptr(), n * n, n, mode, ctas, smem, st) == 0
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        nbytes = n ** 3 * 4 * (1 if mode == 2 else 2)
        print(f"mode {mode} ({name}), <= {ctas:3d} CTAs/SM: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s")
