"""The C-ABI library loads and exports every symbol include/mdvc.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mdvc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mdvc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from mdvcontainment_b200 import _lib, build
    build.build_library()
    names = declared_symbols()
    assert len(names) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in mdvc.h but not exported by libmdvc.so"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    bound = _lib.load()
    assert bound.mdvc_version() >= 100
    assert bound.mdvc_row_pitch_words(1) == 4 and bound.mdvc_row_pitch_words(1024) == 32 and bound.mdvc_row_pitch_words(129) == 8


def test_workspace_query_and_argument_checks_without_gpu():
    from mdvcontainment_b200 import _lib
    lib = _lib.load()
    caps = _lib.FrameCaps(1000, 100, 64, 64)
    assert lib.mdvc_frame_workspace_bytes(8, 8, 8, ctypes.byref(caps)) > 0
    bad = _lib.FrameCaps(1000, 100, 100, 64)            # slots must be a power of two
    assert lib.mdvc_frame_workspace_bytes(8, 8, 8, ctypes.byref(bad)) == 0
    assert lib.mdvc_morph(None, None, None, 4, 4, 4, 4, b"de", None) == _lib.ERR_BADARG
    assert lib.mdvc_pack_bool(None, 4, 4, 4, None, 4, None) == _lib.ERR_BADARG


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mdvcontainment_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "voxel_oracle" not in src and "refshim" not in src and "oracle/" not in src, f
