"""Makes the reference's *Python* package importable in this container (test infrastructure only).

/root/reference is read-only and exists only here, never on the GPU box, so everything that uses this module is a
``not gpu`` test or a golden-vector generator.  ``import projectq`` pulls matplotlib (absent here) through
backends/_circuits/_plot.py, and ``projectq.backends._sim._simulator`` wants the compiled ``_cppsim``; both are
satisfied with in-memory stand-ins: a stub matplotlib, and the reference extension built by oracle/Makefile."""
import importlib.util
import os
import sys
import types

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available():
    return os.path.isdir(os.path.join(REF, "projectq"))


def _stub_matplotlib():
    if "matplotlib" in sys.modules:
        return
    mpl = types.ModuleType("matplotlib")
    for sub, names in {"pyplot": [], "collections": ["LineCollection", "PatchCollection"], "lines": ["Line2D"],
                       "patches": ["Circle", "Arc", "Rectangle"]}.items():
        m = types.ModuleType("matplotlib." + sub)
        for n in names:
            setattr(m, n, type(n, (), {}))
        setattr(mpl, sub, m)
        sys.modules["matplotlib." + sub] = m
    sys.modules["matplotlib"] = mpl


def import_projectq():
    """import the reference package from /root/reference with the compiled reference _cppsim attached"""
    if not available():
        return None
    _stub_matplotlib()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import glob

    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "_cppsim*.so"))
    if hits and "projectq.backends._sim._cppsim" not in sys.modules:
        spec = importlib.util.spec_from_file_location("projectq.backends._sim._cppsim", hits[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules["projectq.backends._sim._cppsim"] = mod
    import projectq
    import projectq.backends._sim as _sim_pkg

    if "projectq.backends._sim._cppsim" in sys.modules:
        _sim_pkg._cppsim = sys.modules["projectq.backends._sim._cppsim"]
    return projectq
