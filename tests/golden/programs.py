"""The ProjectQ programs behind the golden fixtures, written once and run twice:

* by ``make_golden.py`` in the build container — reference engine + reference C++ simulator, every native-seam call recorded;
* by ``tests/test_gpu_dropin.py`` on the B200 box — ``MainEngine(projectq_b200.Simulator(...))`` on the CUDA backend; the
  measured bits, energies and probabilities must reproduce the fixtures.

Every function takes ``make_sim(gate_fusion, rnd_seed)`` returning the backend engine to hand to ``MainEngine`` and
returns ``(results, sim, qureg, eng)``.  Needs the ``projectq`` package (tests/refenv.py makes it importable).
"""
import numpy as np


def qft20(make_sim):
    """BASELINE config 1 as written: MainEngine(Simulator(gate_fusion=True, rnd_seed=1)), default engine list"""
    from projectq import MainEngine
    from projectq.ops import QFT, All, Measure, Ry

    sim = make_sim(True, 1)
    eng = MainEngine(sim)  # default engine list
    q = eng.allocate_qureg(20)
    rng = np.random.default_rng(20)
    for i in range(20):  # seeded product-state preparation so that the QFT output has structure
        Ry(float(rng.uniform(0, np.pi))) | q[i]
    QFT | q
    eng.flush()
    yield "state", sim, q, rng
    All(Measure) | q
    eng.flush()
    yield "done", {"measured_bits": [int(b) for b in q]}


def shor(make_sim, N=4087, a=7, seed=3):
    """BASELINE config 3: the body of examples/shor.py run_shor (:31-86) with emulation on (the InstructionFilter lets
    the constant-math gates through to the simulator)"""
    import projectq.libs.math
    import projectq.setups.decompositions
    from projectq import MainEngine
    from projectq.cengines import AutoReplacer, DecompositionRuleSet, InstructionFilter, LocalOptimizer, TagRemover
    from projectq.libs.math import AddConstant, AddConstantModN, MultiplyByConstantModN
    from projectq.meta import Control
    from projectq.ops import QFT, All, BasicMathGate, H, Measure, R, Swap, X, get_inverse

    sim = make_sim(True, seed)

    def high_level_gates(eng, cmd):
        g = cmd.gate
        if g == QFT or get_inverse(g) == QFT or g == Swap:
            return True
        if isinstance(g, BasicMathGate):
            return isinstance(g, (AddConstant, AddConstantModN, MultiplyByConstantModN))
        return eng.next_engine.is_available(cmd)

    rule_set = DecompositionRuleSet(modules=[projectq.libs.math, projectq.setups.decompositions])
    engines = [AutoReplacer(rule_set), InstructionFilter(high_level_gates), TagRemover(), LocalOptimizer(3),
               AutoReplacer(rule_set), TagRemover(), LocalOptimizer(3)]
    eng = MainEngine(sim, engines)
    n = int(np.ceil(np.log2(N)))
    x = eng.allocate_qureg(n)
    X | x[0]
    measurements = [0] * (2 * n)
    ctrl_qubit = eng.allocate_qubit()
    for k in range(2 * n):
        current_a = pow(a, 1 << (2 * n - 1 - k), N)
        H | ctrl_qubit
        with Control(eng, ctrl_qubit):
            MultiplyByConstantModN(current_a, N) | x
        for i in range(k):
            if measurements[i]:
                R(-np.pi / (1 << (k - i))) | ctrl_qubit
        H | ctrl_qubit
        Measure | ctrl_qubit
        eng.flush()
        measurements[k] = int(ctrl_qubit)
        if measurements[k]:
            X | ctrl_qubit
    All(Measure) | x
    eng.flush()
    yield "done", {"measurements": measurements, "x_bits": [int(b) for b in x]}


def tfim(make_sim, n=12, seed=4):
    """BASELINE config 4 at 12 qubits: Ry prep, 2 x [TimeEvolution(0.3, H), <H>]"""
    from projectq import MainEngine
    from projectq.ops import All, Measure, QubitOperator, Ry, TimeEvolution

    sim = make_sim(True, seed)
    eng = MainEngine(sim, [])
    q = eng.allocate_qureg(n)
    rng = np.random.default_rng(seed)
    for i in range(n):
        Ry(float(rng.uniform(0, np.pi))) | q[i]
    Hop = QubitOperator(())
    Hop *= 0.0
    for i in range(n - 1):
        Hop += QubitOperator("Z%d Z%d" % (i, i + 1), -1.0)
    for i in range(n):
        Hop += QubitOperator("X%d" % i, -0.7)
    eng.flush()
    energies = [sim.get_expectation_value(Hop, q)]
    for it in range(2):
        TimeEvolution(0.3, Hop) | q
        eng.flush()
        energies.append(sim.get_expectation_value(Hop, q))
    yield "state", sim, q, rng
    p = sim.get_probability("010", q[:3])
    All(Measure) | q
    eng.flush()
    yield "done", {"energies": energies, "p010": p, "bits": [int(b) for b in q]}


def brickwork20(make_sim, seed=6):
    """BASELINE config 2 at 20 qubits: the generator of projectq_b200/workloads.py as ProjectQ gates, gate_fusion=True,
    empty engine list"""
    from projectq import MainEngine
    from projectq.ops import CNOT, CZ, All, Measure, Rx, Ry, Rz

    n, depth = 20, 20
    sim = make_sim(True, seed)
    eng = MainEngine(sim, [])
    q = eng.allocate_qureg(n)
    rng = np.random.default_rng(2026)
    for d in range(depth):
        for i in range(n):
            kind = int(rng.integers(0, 3))
            th = float(rng.uniform(0, 2 * np.pi))
            (Rx, Ry, Rz)[kind](th) | q[i]
        for i in range(d % 2, n - 1, 2):
            if int(rng.integers(0, 2)) == 0:
                CNOT | (q[i], q[i + 1])
            else:
                CZ | (q[i], q[i + 1])
    eng.flush()
    yield "state", sim, q, rng
    p = sim.get_probability("0110", q[3:7])
    All(Measure) | q
    eng.flush()
    yield "done", {"p0110": p, "bits": [int(b) for b in q]}


PROGRAMS = {"qft20": qft20, "shor4087": shor, "tfim12": tfim, "brickwork20": brickwork20}
SAMPLES = {"qft20": 256, "tfim12": 256, "brickwork20": 512}


def sample_indices(n_amps, n_samples, rng):
    """the seeded sample of basis indices whose amplitudes a fixture records after the "state" checkpoint"""
    return sorted(int(i) for i in rng.choice(n_amps, size=min(n_samples, n_amps), replace=False))
