"""Makes the reference's *Python* package importable (test infrastructure only).

/root/reference is read-only and exists only in the build container; oracle/Makefile stages a copy of the package under
the git-ignored baseline/_ref/, which travels to the GPU box with the other built artefacts, so the drop-in tests
(`MainEngine(projectq_b200.Simulator(...))`, the reference's own test-suite on the CUDA backend) can run there.
``import projectq`` pulls matplotlib (absent here) through
backends/_circuits/_plot.py, and ``projectq.backends._sim._simulator`` wants the compiled ``_cppsim``; both are
satisfied with in-memory stand-ins: a stub matplotlib, and the reference extension built by oracle/Makefile."""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
if not os.path.isdir(os.path.join(REF, "projectq")):
    REF = os.path.join(ROOT, "baseline", "_ref")


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


def cuda_native_module():
    """a stand-in for the reference's `_cppsim` extension module whose `Simulator` is the CUDA backend — the one-line
    rebinding of INTEGRATION.md (`from projectq_b200.backend import SimulatorBackend as Simulator`)"""
    from projectq_b200.backend import SimulatorBackend

    mod = types.ModuleType("projectq.backends._sim._cppsim")
    mod.Simulator = SimulatorBackend
    mod.__doc__ = "projectq_b200 CUDA backend bound in place of the reference's _cppsim"
    return mod


def import_projectq(native="reference"):
    """import the reference package with a native simulator module attached as projectq.backends._sim._cppsim:
    native="reference" -> the compiled reference extension (oracle/_ref/_cppsim*.so); native="cuda" -> the CUDA backend"""
    if not available():
        return None
    _stub_matplotlib()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import glob

    name = "projectq.backends._sim._cppsim"
    if native == "cuda":
        if name not in sys.modules or getattr(sys.modules[name], "__file__", None):
            sys.modules[name] = cuda_native_module()
    else:
        hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "_cppsim*.so"))
        if hits and name not in sys.modules:
            spec = importlib.util.spec_from_file_location(name, hits[0])
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            sys.modules[name] = mod
    import projectq
    import projectq.backends._sim as _sim_pkg

    if name in sys.modules:
        _sim_pkg._cppsim = sys.modules[name]
    return projectq
