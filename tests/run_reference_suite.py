"""Runs the reference's own Simulator tests against this repo's drop-in pieces.

Usage: python tests/run_reference_suite.py [pytest args]

Two switches (environment):
  PQB_NATIVE = 0 (default)  the native object behind the engine is the compiled reference `_cppsim` (no GPU needed):
                            exercises every line of our command dispatch and user API against the reference's expectations
             = 1            `projectq.backends._sim._cppsim.Simulator` is the CUDA backend (the rebinding of INTEGRATION.md):
                            the tests' own `cpp_simulator` fixture (_simulator_test.py:80-93) then puts CUDA under the engine
  PQB_ENGINE = ours (default)   `projectq.backends.Simulator` is projectq_b200.Simulator (our engine class)
             = reference        the reference's own Python engine class stays; only the native seam is swapped

The reference package comes from /root/reference in the build container and from the staged copy under baseline/_ref/
on the GPU box (tests/refenv.py).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from tests import refenv  # noqa: E402

NATIVE = os.environ.get("PQB_NATIVE") == "1"
ENGINE = os.environ.get("PQB_ENGINE", "ours")

if refenv.import_projectq("cuda" if NATIVE else "reference") is None:
    print("reference not available")
    sys.exit(0)

import projectq.backends  # noqa: E402
import projectq.backends._sim  # noqa: E402
import projectq.backends._sim._cppsim as native_mod  # noqa: E402

if ENGINE == "ours":
    import projectq_b200._simulator as ours  # noqa: E402

    ours.SimulatorBackend = native_mod.Simulator
    projectq.backends.Simulator = ours.Simulator
    projectq.backends._sim.Simulator = ours.Simulator
    # the unitary backend: the reference's test module does `from ._unitary import UnitarySimulator`
    import projectq.backends._unitary as ref_unitary  # noqa: E402

    import projectq_b200._unitary as ours_unitary  # noqa: E402

    ours_unitary.SimulatorBackend = native_mod.Simulator
    ref_unitary.UnitarySimulator = ours_unitary.UnitarySimulator
    projectq.backends.UnitarySimulator = ours_unitary.UnitarySimulator
else:
    import projectq.backends._sim._simulator as ref_engine  # noqa: E402

    ref_engine.SimulatorBackend = native_mod.Simulator
    ref_engine.FALLBACK_TO_PYSIM = False

import pytest  # noqa: E402

REF = refenv.REF
default = [os.path.join(REF, "projectq/backends/_sim/_simulator_test.py"), os.path.join(REF, "projectq/tests/_factoring_test.py")]
if ENGINE == "ours":
    default.append(os.path.join(REF, "projectq/backends/_unitary_test.py"))
args = sys.argv[1:] or default
print("engine=%s native=%s (%s)" % (ENGINE, "cuda" if NATIVE else "reference", native_mod.Simulator), flush=True)
sys.exit(pytest.main(["-q", "-p", "no:cacheprovider", "-p", "no:warnings", "--rootdir", "/tmp"] + args))
