"""Build recipe for ``oracle/_ref`` -- TEST INFRASTRUCTURE ONLY.

Compiles the reference's three Cython kernels *from where they lie* under
``/root/reference/mdvcontainment`` (find_label_contacts.pyx, find_bridges.pyx,
atoms_voxels_mapping.pyx; the recipe is the one of the reference's
``setup.py:15-16``: cythonize + numpy include dir, no extra flags) and writes
ONLY the resulting shared objects into ``oracle/_ref/``.  No reference source
is copied into this repository: the generated C lives in a temp dir that is
removed afterwards.  ``oracle/_ref/`` is git-ignored but travels to the GPU box.

The compiled modules are used (a) by ``oracle/refshim.py`` to run the whole,
unmodified reference package in this container (golden-vector generation) and
(b) by ``bench.py --impl reference`` / ``cpu_baseline`` as the reference's own
native code for the contact/bridge/mapping stages.

Run:  python oracle/build_ref.py            (no-op when /root/reference is absent)
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PKG = os.environ.get("MDVC_REFERENCE", "/root/reference") + "/mdvcontainment"
OUT = os.path.join(HERE, "_ref")
MODULES = ("find_label_contacts", "find_bridges", "atoms_voxels_mapping")


def ext_suffix():
    return sysconfig.get_config_var("EXT_SUFFIX") or ".so"


def built():
    return all(os.path.exists(os.path.join(OUT, m + ext_suffix())) for m in MODULES)


def build(force=False, verbose=True):
    """Compile the reference Cython kernels into oracle/_ref. Returns True if available."""
    if built() and not force:
        return True
    if not os.path.isdir(REF_PKG):
        if verbose:
            print(f"[oracle/_ref] {REF_PKG} not present; nothing to build", file=sys.stderr)
        return built()
    import numpy
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="mdvc_ref_build_")
    try:
        inc = [sysconfig.get_paths()["include"], numpy.get_include()]
        for m in MODULES:
            pyx = os.path.join(REF_PKG, m + ".pyx")
            c_file = os.path.join(tmp, m + ".c")
            subprocess.check_call(
                [sys.executable, "-m", "cython", "-3", pyx, "-o", c_file],
                stdout=subprocess.DEVNULL if not verbose else None)
            so = os.path.join(OUT, m + ext_suffix())
            cmd = ["gcc", "-O2", "-shared", "-fPIC", "-fwrapv", "-w",
                   "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"]
            for i in inc:
                cmd += ["-I", i]
            cmd += [c_file, "-o", so]
            subprocess.check_call(cmd)
            if verbose:
                print(f"[oracle/_ref] built {so}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    sys.exit(0 if ok or not os.path.isdir(REF_PKG) else 1)
