"""Device-side building blocks: PyTorch tensors as buffers, every computation through the C-ABI.

PyTorch is plumbing here (allocation, streams, host<->device copies); no tensor op on the hot
path is a torch kernel.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import FrameCaps, FrameHeader, MdvcError, check


def _require_cuda():
    if not torch.cuda.is_available():
        raise MdvcError("mdvcontainment_b200 needs a CUDA device (B200); there is no CPU fallback.")


# torch.cuda.current_stream() resolves its device argument through is_available() / NVML / os.getenv on every call
# (14 us each; thirteen calls per frame were a third of the host time of a 512^3 trajectory frame).  The helpers below
# go straight to the C bindings the public functions wrap, and fall back to the public functions if a torch build
# lacks them.
_RAW = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_CUR = getattr(torch._C, "_cuda_getCurrentStream", None)
_SET = getattr(torch._C, "_cuda_setStream", None)


def _stream():
    """The current stream of the current device as a ctypes handle for the C-ABI."""
    if _RAW is not None:
        return ctypes.c_void_p(_RAW(torch._C._cuda_getDevice()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def current_stream(device_index: int | None = None) -> torch.cuda.Stream:
    """torch.cuda.current_stream() without the device bookkeeping."""
    if _CUR is None:
        return torch.cuda.current_stream(device_index)
    sd = _CUR(torch._C._cuda_getDevice() if device_index is None else device_index)
    return torch.cuda.Stream(stream_id=sd[0], device_index=sd[1], device_type=sd[2])


class on_stream:
    """``with on_stream(s):`` = ``with torch.cuda.stream(s):`` for a stream of the current device, two C calls each way."""
    __slots__ = ("s", "prev")

    def __init__(self, s: torch.cuda.Stream):
        self.s = s

    def __enter__(self):
        s = self.s
        if _CUR is None or _SET is None:
            self.prev = torch.cuda.current_stream(s.device_index)
            torch.cuda.set_stream(s)
        else:
            self.prev = _CUR(s.device_index)
            _SET(stream_id=s.stream_id, device_index=s.device_index, device_type=s.device_type)
        return s

    def __exit__(self, *exc):
        p = self.prev
        if isinstance(p, tuple):
            _SET(stream_id=p[0], device_index=p[1], device_type=p[2])
        else:
            torch.cuda.set_stream(p)
        return False


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def pitch_words(Z: int) -> int:
    return ((((Z + 31) // 32) + 3) // 4) * 4


class BitGrid:
    """Bit-packed occupancy grid on the device: int32 tensor [X, Y, pitch] (bit k of word w = z 32w+k)."""

    __slots__ = ("shape", "pitch", "words")

    def __init__(self, shape, words=None, device=None):
        _require_cuda()
        self.shape = tuple(int(s) for s in shape)
        if len(self.shape) != 3 or min(self.shape) <= 0:
            raise ValueError(f"grid must be a non-empty 3D array, got shape {self.shape}")
        if int(np.prod(self.shape, dtype=np.int64)) >= 2 ** 31:
            raise ValueError("grids with >= 2^31 voxels need slab decomposition")
        if self.shape[2] >= 1 << 24:
            raise ValueError("rows of >= 2^24 voxels are not supported (run z-starts are packed into 24 bits)")
        self.pitch = pitch_words(self.shape[2])
        if words is None:
            words = torch.zeros((self.shape[0], self.shape[1], self.pitch), dtype=torch.int32,
                                device=device or "cuda")
        self.words = words

    @property
    def n_voxels(self):
        return self.shape[0] * self.shape[1] * self.shape[2]

    @classmethod
    def from_bool(cls, grid) -> "BitGrid":
        """Pack a bool/uint8 array (NumPy or torch, host or device)."""
        _require_cuda()
        if isinstance(grid, np.ndarray):
            t = to_device(np.ascontiguousarray(grid).view(np.uint8) if grid.dtype == np.bool_
                          else np.ascontiguousarray(grid != 0).view(np.uint8))
        else:
            t = grid.to(device="cuda")
            t = (t if t.dtype == torch.bool else (t != 0)).contiguous().view(torch.uint8)
        out = cls(t.shape)
        X, Y, Z = out.shape
        check(_lib.load().mdvc_pack_bool(_ptr(t), X, Y, Z, _ptr(out.words), out.pitch, _stream()), "mdvc_pack_bool")
        return out

    def to_bool(self) -> torch.Tensor:
        X, Y, Z = self.shape
        out = torch.empty(self.shape, dtype=torch.uint8, device=self.words.device)
        check(_lib.load().mdvc_unpack_bool(_ptr(self.words), X, Y, Z, self.pitch, _ptr(out), _stream()), "mdvc_unpack_bool")
        return out.view(torch.bool)

    def count(self) -> int:
        X, Y, Z = self.shape
        c = torch.zeros(1, dtype=torch.int64, device=self.words.device)
        check(_lib.load().mdvc_popcount(_ptr(self.words), X, Y, Z, self.pitch, _ptr(c), _stream()), "mdvc_popcount")
        return int(c.item())


def morph(bits: BitGrid, ops: str) -> BitGrid:
    """Periodic 3x3x3 dilation/erosion sequence on the bit grid (input untouched)."""
    out = BitGrid(bits.shape, device=bits.words.device)
    X, Y, Z = bits.shape
    tmp = torch.empty_like(bits.words) if len(ops) > 1 else None
    rc = _lib.load().mdvc_morph(_ptr(bits.words), _ptr(out.words), _ptr(tmp), X, Y, Z, bits.pitch,
                                ops.encode(), _stream())
    check(rc, "mdvc_morph")
    return out


def bin_atoms(pos: torch.Tensor, T, invT, nbox, bits: BitGrid | None, want_voxels=False, want_positions=False):
    """A1 on the device.  pos: float32 [n,3] cuda tensor.  Returns (voxels int32 [n,3] | None, new positions | None)."""
    n = int(pos.shape[0])
    Tc = np.ascontiguousarray(T, dtype=np.float64)
    iTc = np.ascontiguousarray(invT, dtype=np.float64)
    nb = np.ascontiguousarray(nbox, dtype=np.int64)
    vox = torch.empty((n, 3), dtype=torch.int32, device=pos.device) if want_voxels else None
    newpos = torch.empty((n, 3), dtype=torch.float32, device=pos.device) if want_positions else None
    rc = _lib.load().mdvc_bin_atoms(_ptr(pos), n, Tc.ctypes.data_as(ctypes.c_void_p), iTc.ctypes.data_as(ctypes.c_void_p),
                                    nb.ctypes.data_as(ctypes.c_void_p), _ptr(bits.words if bits is not None else None),
                                    bits.pitch if bits is not None else 0, _ptr(vox), _ptr(newpos), _stream())
    check(rc, "mdvc_bin_atoms")
    return vox, newpos


def bin_scratch(n: int, device) -> torch.Tensor:
    """Scratch for ``bin_grid`` (bucketed bead keys), reusable for any bead count <= n."""
    return torch.empty(max(int(_lib.load().mdvc_bin_scratch_bytes(int(n))), 16), dtype=torch.uint8, device=device)


def bin_grid(pos: torch.Tensor, T, invT, nbox, bits: BitGrid, scratch: torch.Tensor | None = None, want_voxels=False,
             want_positions=False):
    """A1 into ``bits``, which this call also clears (no separate memset): beads bucketed by x-plane range, the grid
    cleared and filled range by range inside the L2 (``mdvc_bin_grid``).  Returns (voxels | None, positions | None)."""
    n = int(pos.shape[0])
    Tc = np.ascontiguousarray(T, dtype=np.float64)
    iTc = np.ascontiguousarray(invT, dtype=np.float64)
    nb = np.ascontiguousarray(nbox, dtype=np.int64)
    need = int(_lib.load().mdvc_bin_scratch_bytes(n))
    if scratch is None or scratch.numel() < need:
        scratch = bin_scratch(n, pos.device)
    vox = torch.empty((n, 3), dtype=torch.int32, device=pos.device) if want_voxels else None
    newpos = torch.empty((n, 3), dtype=torch.float32, device=pos.device) if want_positions else None
    rc = _lib.load().mdvc_bin_grid(_ptr(pos), n, Tc.ctypes.data_as(ctypes.c_void_p), iTc.ctypes.data_as(ctypes.c_void_p),
                                   nb.ctypes.data_as(ctypes.c_void_p), _ptr(bits.words), bits.pitch, _ptr(vox), _ptr(newpos),
                                   _ptr(scratch), int(scratch.numel()), _stream())
    check(rc, "mdvc_bin_grid")
    return vox, newpos


# ---------------------------------------------------------------------------------------------
# host staging memory
# ---------------------------------------------------------------------------------------------
class _HostBlock:
    """Owner of one mdvc_host_alloc block (freed with the last tensor that views it)."""

    def __init__(self, nbytes, write_combined):
        self.ptr = ctypes.c_void_p()
        check(_lib.load().mdvc_host_alloc(int(nbytes), int(bool(write_combined)), ctypes.byref(self.ptr)), "mdvc_host_alloc")
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if self.ptr:
                _lib.load().mdvc_host_free(self.ptr)
                self.ptr = ctypes.c_void_p()
        except Exception:                               # noqa: BLE001 - interpreter shutdown
            pass


def staging_empty(shape, dtype=torch.float32, write_combined=True) -> torch.Tensor:
    """Page-locked host tensor for H2D staging.  ``write_combined``: fill it front to back, never read it on the CPU."""
    _require_cuda()
    shape = tuple(int(v) for v in shape)
    n = int(np.prod(shape)) if shape else 1
    item = torch.empty((), dtype=dtype).element_size()
    block = _HostBlock(max(n * item, 1), write_combined)
    buf = (ctypes.c_uint8 * block.nbytes).from_address(block.ptr.value)
    buf._mdvc_block = block                            # tensor storage -> NumPy array -> ctypes buffer -> allocation
    arr = np.frombuffer(buf, dtype=np.uint8)
    return torch.from_numpy(arr)[: n * item].view(dtype).reshape(shape)


# ---------------------------------------------------------------------------------------------
# device tensor -> NumPy array
# ---------------------------------------------------------------------------------------------
_PINNED = []          # two reusable pinned staging buffers (uint8)
_PINNED_BYTES = 64 << 20


def to_numpy(t: torch.Tensor) -> np.ndarray:
    """``t.cpu().numpy()`` for large device tensors at PCIe speed: the copy into freshly allocated pageable memory
    runs at ~2 GB/s (4 GB components grid: 2.0 s on the B200 box); staging 64 MB chunks through two pinned buffers
    while the previous chunk is copied out by the host gives ~17 GB/s (0.24 s)."""
    if not t.is_cuda or t.numel() * t.element_size() < (8 << 20):
        return t.cpu().numpy()
    src = t.contiguous().reshape(-1)
    as_bool = src.dtype == torch.bool
    if as_bool:
        src = src.view(torch.uint8)
    if not _PINNED:
        _PINNED.extend(torch.empty(_PINNED_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(2))
    item = src.element_size()
    ce = _PINNED_BYTES // item
    out = np.empty(src.numel(), dtype=np.uint8 if as_bool else torch.empty((), dtype=src.dtype).numpy().dtype)
    bufs = [b.view(src.dtype)[:ce] for b in _PINNED]
    evs = [torch.cuda.Event(), torch.cuda.Event()]
    n = src.numel()
    starts = list(range(0, n, ce))
    dst = torch.from_numpy(out)

    def drain(i):
        s0 = starts[i]
        e0 = min(s0 + ce, n)
        evs[i % 2].synchronize()
        dst[s0:e0].copy_(bufs[i % 2][:e0 - s0])

    for i, s0 in enumerate(starts):
        e0 = min(s0 + ce, n)
        bufs[i % 2][:e0 - s0].copy_(src[s0:e0], non_blocking=True)
        evs[i % 2].record()
        if i > 0:
            drain(i - 1)
    drain(len(starts) - 1)
    if as_bool:
        out = out.view(np.bool_)
    return out.reshape(tuple(t.shape))


def to_device(a: np.ndarray, device="cuda") -> torch.Tensor:
    """``torch.from_numpy(a).cuda()`` for large host arrays: pageable memory goes through the same two pinned
    staging buffers (host copy of chunk i+1 overlaps the DMA of chunk i)."""
    a = np.ascontiguousarray(a)
    if a.nbytes < (8 << 20):
        return torch.from_numpy(a).to(device)
    as_bool = a.dtype == np.bool_
    src = torch.from_numpy(a.view(np.uint8) if as_bool else a).reshape(-1)
    if not _PINNED:
        _PINNED.extend(torch.empty(_PINNED_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(2))
    ce = _PINNED_BYTES // src.element_size()
    out = torch.empty(src.numel(), dtype=src.dtype, device=device)
    bufs = [b.view(src.dtype)[:ce] for b in _PINNED]
    evs = [torch.cuda.Event(), torch.cuda.Event()]
    n = src.numel()
    for i, s0 in enumerate(range(0, n, ce)):
        e0 = min(s0 + ce, n)
        if i >= 2:
            evs[i % 2].synchronize()                  # the DMA that last read this staging buffer is done
        bufs[i % 2][:e0 - s0].copy_(src[s0:e0])
        out[s0:e0].copy_(bufs[i % 2][:e0 - s0], non_blocking=True)
        evs[i % 2].record()
    for e in evs:
        e.synchronize()                                # the staging buffers are shared: leave nothing in flight
    out = out.view(torch.bool) if as_bool else out
    return out.reshape(a.shape)


# ---------------------------------------------------------------------------------------------
# frame pipeline
# ---------------------------------------------------------------------------------------------
def _pow2_at_least(v):
    p = 64
    while p < v:
        p <<= 1
    return p


class FramePlan:
    """Workspace + capacities for the run-based pipeline of one grid shape (reusable across frames)."""

    def __init__(self, shape, max_runs=None, max_labels=None, contact_slots=None, bridge_slots=None, device="cuda"):
        _require_cuda()
        self.shape = tuple(int(s) for s in shape)
        X, Y, Z = self.shape
        N = X * Y * Z
        if max_runs is None:            # smooth structures have a few runs per row; noise is retried bigger
            max_runs = max(4096, min(N, 24 * X * Y + 4096))
        if max_labels is None:
            max_labels = max(4096, min(N // 8 + 3, max_runs // 4))
        self.caps = FrameCaps(int(max_runs), int(max_labels), _pow2_at_least(contact_slots or 1 << 14),
                              _pow2_at_least(bridge_slots or 1 << 14))
        self.lib = _lib.load()
        nbytes = self.lib.mdvc_frame_workspace_bytes(X, Y, Z, ctypes.byref(self.caps))
        if nbytes == 0:
            raise MdvcError("mdvc_frame_workspace_bytes rejected the capacities")
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.nbytes = nbytes
        self.header = None

    # capacities large enough for the sizes a failed attempt reported
    def grown(self, header: FrameHeader) -> "FramePlan":
        c = self.caps
        runs = max(c.max_runs, int(header.total_runs * 1.25) + 1024) if header.flags & _lib.OVERFLOW_RUNS else c.max_runs
        labels = c.max_labels
        if header.flags & _lib.OVERFLOW_LABELS:
            labels = max(c.max_labels * 4, header.n_true + header.n_false + 1024)
        cs = c.contact_slots * 8 if header.flags & _lib.OVERFLOW_CONTACTS else c.contact_slots
        bs = c.bridge_slots * 8 if header.flags & _lib.OVERFLOW_BRIDGES else c.bridge_slots
        return FramePlan(self.shape, runs, labels, cs, bs, device=self.ws.device)

    def _args(self):
        return _ptr(self.ws), ctypes.c_size_t(self.nbytes), ctypes.byref(self.caps)

    def label(self, bits: BitGrid, rows_counted: bool = False):
        """Stage 1 (runs, linking, numbering).  ``rows_counted``: the per-row run counts are already in
        ``row_counts_ptr()`` (written by ``mdvc_morph_counted``), so the first pass over the grid is skipped."""
        X, Y, Z = self.shape
        check(self.lib.mdvc_frame_label_counted(_ptr(bits.words), X, Y, Z, bits.pitch, *self._args(), int(rows_counted),
                                                _stream()), "mdvc_frame_label_counted")

    def row_counts_ptr(self) -> int:
        X, Y, Z = self.shape
        return self.lib.mdvc_frame_row_counts_ptr(_ptr(self.ws), X, Y, Z, ctypes.byref(self.caps))

    def pairs(self, bits: BitGrid, periodic=True):
        X, Y, Z = self.shape
        check(self.lib.mdvc_frame_pairs(_ptr(bits.words), X, Y, Z, bits.pitch, *self._args(), int(periodic), _stream()),
              "mdvc_frame_pairs")

    def components(self, periodic=True):
        X, Y, Z = self.shape
        check(self.lib.mdvc_frame_components(X, Y, Z, *self._args(), int(periodic), _stream()), "mdvc_frame_components")

    def set_lut(self, lut_dev: torch.Tensor):
        X, Y, Z = self.shape
        check(self.lib.mdvc_frame_set_lut(X, Y, Z, *self._args(), _ptr(lut_dev), int(lut_dev.numel()), _stream()),
              "mdvc_frame_set_lut")

    def expand(self, bits: BitGrid, labels: torch.Tensor | None, components: torch.Tensor | None):
        X, Y, Z = self.shape
        check(self.lib.mdvc_frame_expand(_ptr(bits.words), X, Y, Z, bits.pitch, *self._args(), _ptr(labels),
                                         _ptr(components), _stream()), "mdvc_frame_expand")

    def atom_values(self, vox: torch.Tensor, which: int) -> torch.Tensor:
        """float64 [n]: label (which=0) or component (which=1) of the voxel of every atom, from the run tables."""
        X, Y, Z = self.shape
        out = torch.empty(vox.shape[0], dtype=torch.float64, device=vox.device)
        check(self.lib.mdvc_frame_atom_values(_ptr(vox), int(vox.shape[0]), X, Y, Z, *self._args(), int(which), _ptr(out),
                                              _stream()), "mdvc_frame_atom_values")
        return out

    def values_at(self, keys: torch.Tensor, which: int) -> torch.Tensor:
        """int32 [n]: label / component at the linear voxel indices ``keys`` (int64/uint64), from the run tables."""
        X, Y, Z = self.shape
        out = torch.empty(keys.shape[0], dtype=torch.int32, device=keys.device)
        check(self.lib.mdvc_frame_values_at(_ptr(keys), int(keys.shape[0]), X, Y, Z, *self._args(), int(which), _ptr(out),
                                            _stream()), "mdvc_frame_values_at")
        return out

    def read_header(self) -> FrameHeader:
        """Synchronises the current stream."""
        h = FrameHeader()
        rc = self.lib.mdvc_frame_read_header(_ptr(self.ws), ctypes.byref(h), _stream())
        check(rc, "mdvc_frame_read_header", allow_overflow=True)
        self.header = h
        return h

    def _view(self, getter, dtype, count, cols=None):
        X, Y, Z = self.shape
        ptr = getter(_ptr(self.ws), X, Y, Z, ctypes.byref(self.caps))
        off = ptr - self.ws.data_ptr()
        itemsize = torch.empty((), dtype=dtype).element_size()
        n = count * (cols or 1)
        t = self.ws[off: off + n * itemsize].view(dtype)
        return t.view(count, cols) if cols else t

    def run_table(self, which: int, count: int) -> np.ndarray:
        """Test access to the run tables (see mdvc_frame_debug_ptr)."""
        X, Y, Z = self.shape
        ptr = self.lib.mdvc_frame_debug_ptr(_ptr(self.ws), X, Y, Z, ctypes.byref(self.caps), which)
        off = ptr - self.ws.data_ptr()
        return self.ws[off: off + 4 * count].view(torch.int32).cpu().numpy()

    def contacts(self) -> np.ndarray:
        h = self.header
        return to_numpy(self._view(self.lib.mdvc_frame_contacts_ptr, torch.int32, h.n_contacts, 2))

    def bridges(self) -> np.ndarray:
        h = self.header
        if h.n_bridges == 0:
            return np.array([], dtype=np.int32)
        return to_numpy(self._view(self.lib.mdvc_frame_bridges_ptr, torch.int32, h.n_bridges, 5))

    def label_volumes(self) -> np.ndarray:
        h = self.header
        return to_numpy(self._view(self.lib.mdvc_frame_label_volumes_ptr, torch.int64, h.n_true + h.n_false + 1))

    def lut(self) -> np.ndarray:
        h = self.header
        return to_numpy(self._view(self.lib.mdvc_frame_lut_ptr, torch.int32, h.n_true + h.n_false + 1))

    def component_counts(self) -> np.ndarray:
        h = self.header
        return self._view(self.lib.mdvc_frame_component_counts_ptr, torch.int64, h.n_comp_pos + h.n_comp_neg + 1).cpu().numpy()


def label_with_retry(bits: BitGrid, plan: FramePlan | None = None, periodic=True, want_pairs=True, max_tries=6):
    """Run label (+pairs) growing the workspace until no capacity overflows. Returns (plan, header)."""
    plan = plan or FramePlan(bits.shape, device=bits.words.device)
    for _ in range(max_tries):
        plan.label(bits)
        if want_pairs:
            plan.pairs(bits, periodic)
        h = plan.read_header()
        if h.flags == 0:
            return plan, h
        plan = plan.grown(h)
    raise MdvcError(f"frame pipeline still overflowing after {max_tries} attempts (flags={h.flags})")


# ---------------------------------------------------------------------------------------------
# seam-level functions on dense int32 grids
# ---------------------------------------------------------------------------------------------
def _as_cuda_i32(grid):
    if isinstance(grid, np.ndarray):
        return to_device(np.ascontiguousarray(grid, dtype=np.int32))
    return grid.to(device="cuda", dtype=torch.int32).contiguous()


def grid_pairs(labels, bridges: bool, n_slots=1 << 12) -> np.ndarray:
    _require_cuda()
    L = _as_cuda_i32(labels)
    X, Y, Z = L.shape
    lib = _lib.load()
    fn = lib.mdvc_grid_bridges if bridges else lib.mdvc_grid_contacts
    cols = 5 if bridges else 2
    while True:
        slots = torch.empty(n_slots, dtype=torch.int64, device=L.device)
        rows = torch.empty((n_slots, cols), dtype=torch.int32, device=L.device)
        meta = torch.zeros(2, dtype=torch.int32, device=L.device)          # [n_rows, flags]
        rc = fn(_ptr(L), X, Y, Z, _ptr(slots), n_slots, _ptr(rows), n_slots, _ptr(meta[0:]), _ptr(meta[1:]), _stream())
        check(rc, "mdvc_grid_bridges" if bridges else "mdvc_grid_contacts")
        n, flags = (int(v) for v in meta.cpu())
        if flags & _lib.OVERFLOW_LABELS:
            raise MdvcError("label magnitude exceeds 2^28")
        if flags == 0:
            break
        n_slots *= 8
    out = rows[:n].cpu().numpy()
    if bridges and n == 0:
        return np.array([], dtype=np.int32)
    return out


def grid_relabel(labels, lut: np.ndarray, lut_lo: int):
    L = _as_cuda_i32(labels)
    lut_d = torch.from_numpy(np.ascontiguousarray(lut, dtype=np.int32)).to(L.device)
    out = torch.empty_like(L)
    check(_lib.load().mdvc_grid_relabel(_ptr(L), L.numel(), _ptr(lut_d), int(lut_lo), int(lut_d.numel()), _ptr(out), _stream()),
          "mdvc_grid_relabel")
    return out


def grid_count(grid, lo: int, n_bins: int) -> np.ndarray:
    G = _as_cuda_i32(grid)
    counts = torch.zeros(n_bins, dtype=torch.int64, device=G.device)
    check(_lib.load().mdvc_grid_count(_ptr(G), G.numel(), int(lo), int(n_bins), _ptr(counts), _stream()), "mdvc_grid_count")
    return counts.cpu().numpy()


def _member_table(values, lo, n):
    m = np.zeros(n, dtype=np.uint8)
    for v in values:
        k = int(v) - lo
        if 0 <= k < n:
            m[k] = 1
    return m


def grid_select(grid: torch.Tensor, values, lo: int, n_values: int, want_mask=False, want_positions=True):
    """np.isin + np.where on the device. Returns (positions int64 [M,3] numpy | None, mask bool tensor | None)."""
    G = _as_cuda_i32(grid)
    X, Y, Z = G.shape
    lib = _lib.load()
    member = torch.from_numpy(_member_table(values, lo, n_values)).to(G.device)
    scratch = torch.empty(lib.mdvc_grid_select_scratch_words(G.numel()), dtype=torch.int32, device=G.device)
    n_out = torch.zeros(1, dtype=torch.int64, device=G.device)
    mask = torch.empty(G.shape, dtype=torch.uint8, device=G.device) if want_mask else None
    # first pass: count (and mask); second pass: scatter into an exactly sized buffer
    check(lib.mdvc_grid_select(_ptr(G), X, Y, Z, _ptr(member), int(lo), int(n_values), _ptr(scratch), None, 0,
                               _ptr(n_out), _ptr(mask), _stream()), "mdvc_grid_select")
    pos = None
    if want_positions:
        m = int(n_out.item())
        pos_t = torch.empty((m, 3), dtype=torch.int64, device=G.device)
        if m:
            check(lib.mdvc_grid_select(_ptr(G), X, Y, Z, _ptr(member), int(lo), int(n_values), _ptr(scratch), _ptr(pos_t),
                                       m, _ptr(n_out), None, _stream()), "mdvc_grid_select")
        pos = to_numpy(pos_t)
    return pos, (mask.view(torch.bool) if mask is not None else None)


def atom_components(vox: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    X, Y, Z = grid.shape
    out = torch.empty(vox.shape[0], dtype=torch.float64, device=vox.device)
    check(_lib.load().mdvc_atom_components(_ptr(vox), vox.shape[0], _ptr(grid), X, Y, Z, _ptr(out), _stream()),
          "mdvc_atom_components")
    return out


class AtomCSR:
    """Atoms stably sorted by the linear index of their voxel (exact voxel -> atoms multimap)."""

    def __init__(self, vox: torch.Tensor, shape):
        self.vox = vox
        self.shape = tuple(int(s) for s in shape)
        n = int(vox.shape[0])
        lib = _lib.load()
        self.order = torch.empty(n, dtype=torch.int32, device=vox.device)
        self.keys = torch.empty(n, dtype=torch.int64, device=vox.device)
        if n:
            scratch = torch.empty(lib.mdvc_atom_sort_scratch_bytes(n), dtype=torch.uint8, device=vox.device)
            X, Y, Z = self.shape
            check(lib.mdvc_atom_sort_by_voxel(_ptr(vox), n, X, Y, Z, _ptr(self.order), _ptr(self.keys), _ptr(scratch),
                                              _stream()), "mdvc_atom_sort_by_voxel")

    def select(self, grid: torch.Tensor, values, lo: int, n_values: int, atom_ix: torch.Tensor | None,
               per_position: bool = False) -> np.ndarray:
        """Atoms (voxel C-order, then atom order) whose voxel value is one of ``values``.  ``grid`` is the dense int32
        grid, or -- ``per_position`` -- one value per entry of this CSR (``FramePlan.values_at(self.keys, ...)``)."""
        n = int(self.order.shape[0])
        if n == 0:
            return np.array([], dtype=np.int64)
        lib = _lib.load()
        member = torch.from_numpy(_member_table(values, lo, n_values)).to(grid.device)
        scratch = torch.empty(lib.mdvc_grid_select_scratch_words(n), dtype=torch.int32, device=grid.device)
        out = torch.empty(n, dtype=torch.int64, device=grid.device)
        n_out = torch.zeros(1, dtype=torch.int64, device=grid.device)
        check(lib.mdvc_atom_select(_ptr(self.order), None if per_position else _ptr(self.keys), n, _ptr(grid), _ptr(member),
                                   int(lo), int(n_values), _ptr(atom_ix), _ptr(scratch), _ptr(out), _ptr(n_out), _stream()),
              "mdvc_atom_select")
        return to_numpy(out[: int(n_out.item())])

    def gather_voxels(self, query_voxels, atom_ix: torch.Tensor | None) -> np.ndarray:
        """Atoms of an explicit voxel list (query order, then atom order) -> int64 atom indices."""
        q = np.ascontiguousarray(np.asarray(query_voxels, dtype=np.int32).reshape(-1, 3))
        m, n = len(q), int(self.order.shape[0])
        if m == 0 or n == 0:
            return np.array([], dtype=np.int64)
        lib = _lib.load()
        dev_ = self.order.device
        qd = torch.from_numpy(q).to(dev_)
        scratch = torch.empty(lib.mdvc_atom_gather_scratch_words(m), dtype=torch.int32, device=dev_)
        n_out = torch.zeros(1, dtype=torch.int64, device=dev_)
        X, Y, Z = self.shape
        args = (_ptr(qd), m, X, Y, Z, _ptr(self.order), _ptr(self.keys), n, _ptr(atom_ix), _ptr(scratch))
        check(lib.mdvc_atom_gather_voxels(*args, None, 0, _ptr(n_out), _stream()), "mdvc_atom_gather_voxels")
        total = int(n_out.item())
        if total == 0:
            return np.array([], dtype=np.int64)
        out = torch.empty(total, dtype=torch.int64, device=dev_)
        check(lib.mdvc_atom_gather_voxels(*args, _ptr(out), total, _ptr(n_out), _stream()), "mdvc_atom_gather_voxels")
        return to_numpy(out)


def tensor_digest(t: torch.Tensor, offset: int = 0, chunk: int = 1 << 26) -> int:
    """Order-dependent 64-bit checksum of an integer / bool tensor on its own device:
    sum(value[i] * ((offset + i) mod 1000003 + 1)) in wrapping int64 arithmetic, so partial digests of the slabs of a
    grid (``offset`` = the slab's first flat index) add up (mod 2^64) to the digest of the whole grid."""
    flat = t.reshape(-1)
    acc = torch.zeros((), dtype=torch.int64, device=flat.device)
    for s in range(0, flat.numel(), chunk):
        part = flat[s:s + chunk].to(torch.int64)
        w = torch.arange(offset + s, offset + s + part.numel(), device=flat.device, dtype=torch.int64) % 1000003 + 1
        acc += (part * w).sum()
    return int(acc.item())
