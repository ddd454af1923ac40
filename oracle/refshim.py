"""Load the UNMODIFIED reference package from /root/reference -- TEST INFRASTRUCTURE ONLY.

Only usable in the build container (``/root/reference`` does not exist on the GPU
box).  Used by ``tests/golden/make_golden.py`` to dump golden vectors and by the
CPU tests that validate ``oracle/voxel_oracle.py`` against the real thing.

The reference imports matplotlib (``containment_main.py:5``, ``graph_logic.py:3``)
and MDAnalysis (``mda_containment.py:4``, ``voxels_to_gro.py:3``) at module import;
neither is installed.  SURVEY.md Appendix C: inject empty matplotlib modules and a
fake ``MDAnalysis`` whose ``core.groups.AtomGroup`` is this repo's duck-typed
``mdvcontainment_b200.universe.AtomGroup``.  The three Cython modules come from
``oracle/_ref`` (built by ``oracle/build_ref.py`` from the sources where they lie).
"""
import importlib
import importlib.machinery
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("MDVC_REFERENCE", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "mdvcontainment")
_REF_SO = os.path.join(HERE, "_ref")

_cached = None


def reference_available() -> bool:
    return os.path.isdir(REF_PKG)


def _install_stubs():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            for sub in ("pyplot", "patches", "colors"):
                m = types.ModuleType("matplotlib." + sub)
                setattr(mpl, sub, m)
                sys.modules["matplotlib." + sub] = m
            sys.modules["matplotlib"] = mpl
    if "MDAnalysis" not in sys.modules:
        try:
            import MDAnalysis  # noqa: F401
        except Exception:
            repo = os.path.dirname(HERE)
            if repo not in sys.path:
                sys.path.insert(0, repo)
            from mdvcontainment_b200 import universe as U
            mda = types.ModuleType("MDAnalysis")
            core = types.ModuleType("MDAnalysis.core")
            groups = types.ModuleType("MDAnalysis.core.groups")
            topo = types.ModuleType("MDAnalysis.core.topologyattrs")
            groups.AtomGroup = U.AtomGroup
            topo.Tempfactors = U.Tempfactors
            core.groups, core.topologyattrs = groups, topo
            mda.core, mda.Universe, mda.AtomGroup = core, U.Universe, U.AtomGroup
            sys.modules.update({"MDAnalysis": mda, "MDAnalysis.core": core,
                                "MDAnalysis.core.groups": groups,
                                "MDAnalysis.core.topologyattrs": topo})


def load_reference():
    """Return the imported reference package ``mdvcontainment`` (module object)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError(f"reference not present at {REF_PKG}")
    sys.path.insert(0, HERE)
    import build_ref
    if not build_ref.build(verbose=False):
        raise RuntimeError("could not build oracle/_ref")
    _install_stubs()
    spec = importlib.util.spec_from_file_location(
        "mdvcontainment", os.path.join(REF_PKG, "__init__.py"),
        submodule_search_locations=[REF_PKG, _REF_SO])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["mdvcontainment"] = mod
    sys.dont_write_bytecode, old = True, sys.dont_write_bytecode
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.dont_write_bytecode = old
    _cached = mod
    return mod


def load_ref_kernels():
    """Import only the compiled reference Cython kernels from ``oracle/_ref``.

    Works on the GPU box too (the .so files travel; no reference sources needed).
    Returns a dict name -> module, or {} if ``oracle/_ref`` was never built."""
    out = {}
    if not os.path.isdir(_REF_SO):
        return out
    for name in ("find_label_contacts", "find_bridges", "atoms_voxels_mapping"):
        full = "mdvcontainment." + name
        if full in sys.modules:
            out[name] = sys.modules[full]
            continue
        cand = [f for f in os.listdir(_REF_SO) if f.startswith(name + ".") and f.endswith(".so")]
        if not cand:
            return {}
        path = os.path.join(_REF_SO, cand[0])
        loader = importlib.machinery.ExtensionFileLoader(name, path)
        spec = importlib.util.spec_from_file_location(name, path, loader=loader)
        try:
            mod = importlib.util.module_from_spec(spec)
            loader.exec_module(mod)
        except Exception:
            return {}
        out[name] = mod
    return out
