"""The Python engine mirror (projectq_b200/_simulator.py) against the reference's own expectations.  CPU container only
(needs the reference's Python package under /root/reference)."""
import os
import re
import subprocess
import sys

import pytest

from tests import refenv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not refenv.available(), reason="/root/reference is not present on this machine")


def test_reference_simulator_suite_passes_on_our_engine():
    """the reference's _simulator_test.py + _factoring_test.py (60 tests) with projectq.backends.Simulator := ours"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_suite.py")], capture_output=True,
                         text=True, timeout=600)
    tail = out.stdout[-3000:]
    m = re.search(r"(\d+) passed", tail)
    assert m, tail
    failed = re.findall(r"FAILED \S+::(\S+)", out.stdout)
    # the one pre-existing failure of the reference itself under NumPy 2 (np.array(list, copy=False) inside the test,
    # _simulator_test.py:562) appears only because the native object here is the reference's list-returning _cppsim
    assert set(failed) <= {"test_simulator_time_evolution[cpp_simulator]"}, tail
    assert int(m.group(1)) >= 59


def test_same_backend_call_sequence_as_reference_engine():
    """QFT + math gates + time evolution through MainEngine: our engine must issue exactly the reference engine's calls"""
    code = r'''
import sys, json
sys.path.insert(0, %r)
sys.dont_write_bytecode = True
from tests import refenv
refenv.import_projectq()
import numpy as np
import projectq.backends._sim._cppsim as ref_cppsim
from projectq.backends import Simulator as RefSimulator
import projectq_b200._simulator as ours
ours.SimulatorBackend = ref_cppsim.Simulator
from projectq import MainEngine
from projectq.ops import QFT, All, Measure, H, X, Rx, CNOT, QubitOperator, TimeEvolution
from projectq.meta import Control
from projectq.libs.math import AddConstant, MultiplyByConstantModN

class Spy:
    def __init__(self, inner, log):
        self._inner, self._log = inner, log
    def __getattr__(self, name):
        f = getattr(self._inner, name)
        def call(*args):
            self._log.append((name, json.dumps(args, default=lambda o: [o.real, o.imag] if isinstance(o, complex) else str(o))))
            return f(*args)
        return call

def program(sim_cls):
    log = []
    sim = sim_cls(gate_fusion=True, rnd_seed=5)
    sim._simulator = Spy(ref_cppsim.Simulator(5), log)
    eng = MainEngine(sim)
    q = eng.allocate_qureg(6)
    X | q[1]; H | q[3]
    QFT | q
    with Control(eng, q[0]):
        Rx(0.3) | q[2]
    eng.flush()
    p = sim.get_probability('01', q[:2])
    anc = eng.allocate_qureg(3)
    X | anc[0]
    eng.flush()
    Hop = QubitOperator('Z0 Z1', 0.7) + QubitOperator('X2', -0.2) + QubitOperator((), 0.1)
    TimeEvolution(0.4, Hop) | q
    eng.flush()
    e = sim.get_expectation_value(Hop, q)
    All(Measure) | q
    All(Measure) | anc
    eng.flush()
    bits = [int(b) for b in q] + [int(b) for b in anc]
    return log, (p, e, bits)

a = program(RefSimulator)
b = program(ours.Simulator)
assert a[1] == b[1], (a[1], b[1])
assert len(a[0]) == len(b[0]), (len(a[0]), len(b[0]))
for x, y in zip(a[0], b[0]):
    assert x == y, (x, y)
print("calls", len(a[0]), "OK")
''' % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "OK" in out.stdout
