"""The full drop-in stack on the B200: ProjectQ's compiler chain -> projectq_b200.Simulator -> pybind shim -> C ABI -> CUDA.

What the reference does to put its C++ simulator under its Python engine (`sim._simulator = CppSim(1)`,
_simulator_test.py:80-93) is done here with the CUDA backend:

* the reference's own test modules (`_simulator_test.py`, `_factoring_test.py`) run unmodified with
  `projectq.backends._sim._cppsim.Simulator` bound to the CUDA backend — once under our engine class and once under the
  reference's own engine class (only the native seam swapped);
* the ProjectQ programs behind the golden fixtures (tests/golden/programs.py: BASELINE configs 1-4 as written, e.g.
  `MainEngine(Simulator(gate_fusion=True, rnd_seed=1))` + QFT + Measure) run on `projectq_b200.Simulator` and must
  reproduce what the reference engine + reference C++ simulator recorded: identical measured bits, energies /
  probabilities / sampled amplitudes within 1e-12;
* `examples/shor.py`'s `run_shor` itself, with emulated and with fully decomposed modular arithmetic.

The reference's Python package is found by tests/refenv.py (staged under baseline/_ref/ by oracle/Makefile).
"""
import importlib.util
import json
import os
import re
import subprocess
import sys
from fractions import Fraction

import numpy as np
import pytest

from tests import refenv

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
TOL = 1e-12


def run_suite(engine):
    env = dict(os.environ, PQB_NATIVE="1", PQB_ENGINE=engine)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_suite.py")], capture_output=True,
                         text=True, timeout=1200, env=env, cwd=ROOT)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    assert "native=cuda" in out.stdout, tail
    m = re.search(r"(?:(\d+) failed, )?(\d+) passed", out.stdout)
    assert m, tail
    return int(m.group(1) or 0), int(m.group(2)), tail


@pytest.mark.parametrize("engine", ["ours", "reference"])
def test_reference_suite_on_cuda(engine):
    """>= 59 of the reference's 60 Simulator/factoring tests pass with CUDA under the engine.  (With the reference's C++
    simulator exactly one fails here — `numpy.array(list, copy=False)` under NumPy 2, _simulator_test.py:562 — and
    cheat() returning an ndarray makes even that one pass.)  Under our engine classes the reference's 8 UnitarySimulator
    tests (_unitary_test.py) run too, against projectq_b200.UnitarySimulator on the CUDA backend."""
    assert refenv.available(), "reference package not staged (baseline/_ref): run `make -C oracle` in the build container"
    failed, passed, tail = run_suite(engine)
    assert passed >= (67 if engine == "ours" else 59) and failed <= 1, tail


@pytest.fixture(scope="module")
def pq():
    assert refenv.available(), "reference package not staged (baseline/_ref): run `make -C oracle` in the build container"
    return refenv.import_projectq("cuda")


def make_cuda_sim(gate_fusion, rnd_seed):
    from projectq_b200 import Simulator

    return Simulator(gate_fusion=gate_fusion, rnd_seed=rnd_seed)


@pytest.mark.parametrize("name", ["qft20", "shor4087", "tfim12", "brickwork20"])
def test_baseline_programs_reproduce_the_reference(pq, name):
    from tests.golden import programs

    with open(os.path.join(GOLDEN, name + ".json")) as f:
        gold = json.load(f)
    checks = [e for e in gold["trace"] if e["m"] == "_check_amplitudes"]
    for step in programs.PROGRAMS[name](make_cuda_sim):
        if step[0] == "state":
            _, sim, _, rng = step
            mapping, vec = sim.cheat()
            idx = programs.sample_indices(len(vec), programs.SAMPLES[name], rng)
            chk = checks.pop(0)
            assert idx == chk["a"][0]
            assert {str(k): int(v) for k, v in dict(mapping).items()} == chk["map"]
            ref = np.array([complex(re, im) for re, im in chk["r"]])
            assert np.max(np.abs(np.asarray(vec)[idx] - ref)) < TOL
        else:
            for key, val in step[1].items():
                if isinstance(gold[key], float) or (isinstance(gold[key], list) and isinstance(gold[key][0], float)):
                    assert np.max(np.abs(np.asarray(val) - np.asarray(gold[key]))) < 1e-11, key
                else:
                    assert val == gold[key], key  # measured bits: identical for the same rnd_seed


def test_unitary_simulator_matches_the_reference_class(pq):
    """projectq_b200.UnitarySimulator (unitary on the GPU) vs the reference's UnitarySimulator (dense NumPy products) on a
    random 6-qubit circuit with controls, multi-qubit gates in scrambled target order and a mid-circuit allocation"""
    from projectq import MainEngine
    from projectq.backends._unitary import UnitarySimulator as RefUnitary
    from projectq.meta import Control
    from projectq.ops import CNOT, H, MatrixGate, Rx, Rz

    from projectq_b200 import UnitarySimulator
    from tests.helpers import rand_unitary

    def program(backend):
        rng = np.random.default_rng(12)
        eng = MainEngine(backend=backend, engine_list=[])
        q = eng.allocate_qureg(5)
        for step in range(30):
            k = int(rng.integers(1, 4))
            qs = [int(x) for x in rng.permutation(5)[: k + 1]]
            gate = MatrixGate(rand_unitary(rng, k))
            if rng.random() < 0.5:
                with Control(eng, q[qs[k]]):
                    gate | tuple(q[i] for i in qs[:k])
            else:
                gate | tuple(q[i] for i in qs[:k])
            if step == 12:
                q = q + eng.allocate_qureg(1)  # U <- 1 (x) U in the middle of the circuit
                H | q[5]
                CNOT | (q[5], q[0])
        Rx(0.3) | q[2]
        Rz(1.1) | q[5]
        eng.flush()
        u = backend.unitary
        from projectq.ops import All, Measure

        All(Measure) | q
        eng.flush()
        return u

    mine, ref = program(UnitarySimulator()), program(RefUnitary())
    assert mine.shape == ref.shape == (64, 64)
    assert np.max(np.abs(mine - ref)) < 1e-12


def load_shor_example():
    path = os.path.join(refenv.REF, "examples", "shor.py")
    spec = importlib.util.spec_from_file_location("reference_examples_shor", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def shor_engine(shor, sim, emulate):
    import projectq.libs.math
    import projectq.setups.decompositions
    from projectq import MainEngine
    from projectq.cengines import AutoReplacer, DecompositionRuleSet, InstructionFilter, LocalOptimizer, TagRemover
    from projectq.libs.math import AddConstant, AddConstantModN, MultiplyByConstantModN
    from projectq.ops import BasicMathGate

    def gate_filter(eng, cmd):
        if emulate and isinstance(cmd.gate, BasicMathGate):
            return isinstance(cmd.gate, (AddConstant, AddConstantModN, MultiplyByConstantModN))
        return shor.high_level_gates(eng, cmd)  # the example's own filter (math gates are decomposed)

    rule_set = DecompositionRuleSet(modules=[projectq.libs.math, projectq.setups.decompositions])
    return MainEngine(sim, [AutoReplacer(rule_set), InstructionFilter(gate_filter), TagRemover(), LocalOptimizer(3),
                            AutoReplacer(rule_set), TagRemover(), LocalOptimizer(3)])


def test_examples_shor_run_shor_emulated(pq):
    """examples/shor.py:31-86 as written, N = 4087, a = 7, emulated modular multiplication: the period candidate must be
    the one the reference's measurement record implies"""
    shor = load_shor_example()
    with open(os.path.join(GOLDEN, "shor4087.json")) as f:
        gold = json.load(f)
    m = gold["measurements"]
    n2 = len(m)
    y = sum(m[n2 - 1 - i] * 1.0 / (1 << (i + 1)) for i in range(n2))
    expected_r = Fraction(y).limit_denominator(4087 - 1).denominator
    eng = shor_engine(shor, make_cuda_sim(True, 3), emulate=True)
    r = shor.run_shor(eng, 4087, 7)
    eng.flush()
    assert r == expected_r


def test_examples_shor_decomposed_arithmetic_matches_reference_simulator(pq):
    """the non-emulated path (libs/math decompositions -> thousands of controlled one- and two-qubit gates on a small
    register): same seed, CUDA backend vs the reference C++ simulator under the same engine class"""
    from tests.conftest import load_ref_cppsim

    ref_mod = load_ref_cppsim()
    assert ref_mod is not None, "oracle/_ref/_cppsim not built"
    shor = load_shor_example()
    results = []
    for native in ("cuda", "reference"):
        sim = make_cuda_sim(True, 11)
        if native == "reference":
            sim._simulator = ref_mod.Simulator(11)
        eng = shor_engine(shor, sim, emulate=False)
        results.append(shor.run_shor(eng, 15, 7))
        eng.flush()
    assert results[0] == results[1]
    assert results[0] in (1, 2, 4)  # 7 has order 4 mod 15
