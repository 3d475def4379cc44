"""Runs the reference's own Simulator tests against *our* engine class (projectq_b200._simulator.Simulator).

Usage: python tests/run_reference_suite.py [pytest args]           (CPU container only: needs /root/reference)

`projectq.backends.Simulator` is replaced by our engine mirror before the reference test modules are imported.  There is
no GPU here, so the native object behind the engine is the compiled reference `_cppsim` (the tests' own fixture swaps
`sim._simulator` the same way, _simulator_test.py:80-93): what this exercises is every line of our command dispatch and
user API against the reference's expectations.  On a GPU box with projectq installed, set PQB_NATIVE=1 to keep the CUDA
backend instead.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from tests import refenv  # noqa: E402

if refenv.import_projectq() is None:
    print("reference not available")
    sys.exit(0)

import projectq.backends  # noqa: E402
import projectq.backends._sim  # noqa: E402
import projectq.backends._sim._cppsim as ref_cppsim  # noqa: E402

import projectq_b200._simulator as ours  # noqa: E402

if os.environ.get("PQB_NATIVE") != "1":
    ours.SimulatorBackend = ref_cppsim.Simulator
projectq.backends.Simulator = ours.Simulator
projectq.backends._sim.Simulator = ours.Simulator

import pytest  # noqa: E402

REF = refenv.REF
default = [os.path.join(REF, "projectq/backends/_sim/_simulator_test.py"), os.path.join(REF, "projectq/tests/_factoring_test.py")]
args = sys.argv[1:] or default
sys.exit(pytest.main(["-q", "-p", "no:cacheprovider", "-p", "no:warnings", "--rootdir", "/tmp"] + args))
