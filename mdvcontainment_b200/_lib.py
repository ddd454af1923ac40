"""ctypes binding of libmdvc.so (the C-ABI declared in include/mdvc.h).

There is no CPU fallback: if the shared object cannot be loaded the import fails loudly.
"""
from __future__ import annotations

import ctypes
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# MDVC_LIB=<path> loads another build of the same ABI (A/B timing of kernel variants; development only)
LIB_PATH = os.environ.get("MDVC_LIB") or os.path.join(PKG, "libmdvc.so")

OK = 0
ERR_BADARG, ERR_CUDA, ERR_WORKSPACE = -1, -2, -3
OVERFLOW_RUNS, OVERFLOW_CONTACTS, OVERFLOW_BRIDGES, OVERFLOW_LABELS, OVERFLOW_ROWS = 1, 2, 4, 8, 16


class MdvcError(RuntimeError):
    pass


class FrameCaps(ctypes.Structure):
    _fields_ = [("max_runs", ctypes.c_uint32), ("max_labels", ctypes.c_uint32),
                ("contact_slots", ctypes.c_uint32), ("bridge_slots", ctypes.c_uint32)]


class FrameHeader(ctypes.Structure):
    _fields_ = [("total_runs", ctypes.c_uint32), ("n_true", ctypes.c_uint32), ("n_false", ctypes.c_uint32),
                ("flags", ctypes.c_uint32), ("n_contacts", ctypes.c_uint32), ("n_bridges", ctypes.c_uint32),
                ("n_comp_pos", ctypes.c_uint32), ("n_comp_neg", ctypes.c_uint32),
                ("reserved", ctypes.c_uint32 * 56)]


_P = ctypes.c_void_p
_I = ctypes.c_int
_LL = ctypes.c_longlong
_U32 = ctypes.c_uint32
_U64 = ctypes.c_uint64
_SZ = ctypes.c_size_t
_CAPS = ctypes.POINTER(FrameCaps)

# name -> (restype, argtypes); must list every symbol include/mdvc.h declares
SIGNATURES = {
    "mdvc_version": (_I, []),
    "mdvc_last_error": (ctypes.c_char_p, []),
    "mdvc_row_pitch_words": (_I, [_I]),
    "mdvc_launch_count": (_U64, []),
    "mdvc_host_alloc": (_I, [_SZ, _I, _P]),
    "mdvc_host_free": (_I, [_P]),
    "mdvc_pack_bool": (_I, [_P, _I, _I, _I, _P, _I, _P]),
    "mdvc_unpack_bool": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "mdvc_popcount": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "mdvc_bin_atoms": (_I, [_P, _LL, _P, _P, _P, _P, _I, _P, _P, _P]),
    "mdvc_bin_scratch_bytes": (_SZ, [_LL]),
    "mdvc_bin_grid": (_I, [_P, _LL, _P, _P, _P, _P, _I, _P, _P, _P, _SZ, _P]),
    "mdvc_morph": (_I, [_P, _P, _P, _I, _I, _I, _I, ctypes.c_char_p, _P]),
    "mdvc_morph_counted": (_I, [_P, _P, _P, _I, _I, _I, _I, ctypes.c_char_p, _P, _P, _P]),
    "mdvc_frame_workspace_bytes": (_SZ, [_I, _I, _I, _CAPS]),
    "mdvc_frame_label": (_I, [_P, _I, _I, _I, _I, _P, _SZ, _CAPS, _P]),
    "mdvc_frame_label_counted": (_I, [_P, _I, _I, _I, _I, _P, _SZ, _CAPS, _I, _P]),
    "mdvc_frame_row_counts_ptr": (_P, [_P, _I, _I, _I, _CAPS]),
    "mdvc_frame_pairs": (_I, [_P, _I, _I, _I, _I, _P, _SZ, _CAPS, _I, _P]),
    "mdvc_frame_components": (_I, [_I, _I, _I, _P, _SZ, _CAPS, _I, _P]),
    "mdvc_frame_set_lut": (_I, [_I, _I, _I, _P, _SZ, _CAPS, _P, _U32, _P]),
    "mdvc_frame_expand": (_I, [_P, _I, _I, _I, _I, _P, _SZ, _CAPS, _P, _P, _P]),
    "mdvc_frame_expand_planes": (_I, [_P, _I, _I, _I, _I, _P, _SZ, _CAPS, _I, _I, _P, _P, _P]),
    "mdvc_frame_relabel_runs": (_I, [_I, _I, _I, _P, _SZ, _CAPS, _P, _U32, _P]),
    "mdvc_frame_atom_values": (_I, [_P, _LL, _I, _I, _I, _P, _SZ, _CAPS, _I, _P, _P]),
    "mdvc_frame_values_at": (_I, [_P, _LL, _I, _I, _I, _P, _SZ, _CAPS, _I, _P, _P]),
    "mdvc_frame_read_header": (_I, [_P, ctypes.POINTER(FrameHeader), _P]),
    "mdvc_frame_contacts_ptr": (_P, [_P, _I, _I, _I, _CAPS]),
    "mdvc_frame_bridges_ptr": (_P, [_P, _I, _I, _I, _CAPS]),
    "mdvc_frame_label_volumes_ptr": (_P, [_P, _I, _I, _I, _CAPS]),
    "mdvc_frame_lut_ptr": (_P, [_P, _I, _I, _I, _CAPS]),
    "mdvc_frame_component_counts_ptr": (_P, [_P, _I, _I, _I, _CAPS]),
    "mdvc_frame_debug_ptr": (_P, [_P, _I, _I, _I, _CAPS, _I]),
    "mdvc_grid_contacts": (_I, [_P, _I, _I, _I, _P, _U32, _P, _U32, _P, _P, _P]),
    "mdvc_grid_bridges": (_I, [_P, _I, _I, _I, _P, _U32, _P, _U32, _P, _P, _P]),
    "mdvc_plane_pairs": (_I, [_P, _P, _I, _I, _I, _I, _P, _U32, _P, _U32, _P, _P, _P]),
    "mdvc_grid_relabel": (_I, [_P, _LL, _P, _I, _I, _P, _P]),
    "mdvc_grid_count": (_I, [_P, _LL, _I, _I, _P, _P]),
    "mdvc_grid_select_scratch_words": (_SZ, [_LL]),
    "mdvc_grid_select": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, _P, _U64, _P, _P, _P]),
    "mdvc_atom_components": (_I, [_P, _LL, _P, _I, _I, _I, _P, _P]),
    "mdvc_atom_sort_scratch_bytes": (_SZ, [_LL]),
    "mdvc_atom_sort_by_voxel": (_I, [_P, _LL, _I, _I, _I, _P, _P, _P, _P]),
    "mdvc_atom_gather_scratch_words": (_SZ, [_LL]),
    "mdvc_atom_gather_voxels": (_I, [_P, _LL, _I, _I, _I, _P, _P, _LL, _P, _P, _P, _U64, _P, _P]),
    "mdvc_atom_select": (_I, [_P, _P, _LL, _P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "mdvc_host_shift_graph_ranks": (_I, [_LL, _P, _P, _P, _LL, _P, _P, _P, _P]),
    "mdvc_host_component_neighbours": (_I, [_LL, _P, _I, _P, _P, _LL, _P, _P, _P]),
    "mdvc_host_adjacency": (_I, [_I, _P, _P, _LL, _P, _P]),
    "mdvc_host_adjacency_unique": (_I, [_I, _P, _P, _LL, _P, _P]),
    "mdvc_host_merge_slabs": (_I, [_I] + [_P] * 18),
}

_lib = None


def load():
    """Load libmdvc.so and bind every declared symbol. Raises MdvcError if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            from . import build
            build.build_library()
        except Exception as exc:  # pragma: no cover
            raise MdvcError(
                f"{LIB_PATH} is missing and could not be built ({exc}); run "
                "`python -m mdvcontainment_b200.build` (needs nvcc). There is no CPU fallback.") from exc
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:
        raise MdvcError(f"cannot load {LIB_PATH}: {exc}. There is no CPU fallback.") from exc
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str, allow_overflow: bool = False) -> int:
    """Turn a status code into an exception (negative) or return it (0 / overflow bits)."""
    if rc == OK:
        return rc
    if rc > 0 and allow_overflow:
        return rc
    lib = load()
    msg = lib.mdvc_last_error().decode(errors="replace")
    names = {ERR_BADARG: "bad argument", ERR_CUDA: "CUDA error", ERR_WORKSPACE: "workspace too small"}
    if rc > 0:
        raise MdvcError(f"{what}: capacity overflow (flags={rc})")
    raise MdvcError(f"{what}: {names.get(rc, rc)}" + (f" [{msg}]" if msg else ""))
