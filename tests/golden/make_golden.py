"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

The reference's Python package is imported from /root/reference and its compiled C++ simulator from oracle/_ref (built
by oracle/Makefile).  Each fixture is the exact sequence of native-seam calls (`_cppsim.Simulator` methods,
reference: _cppsim.cpp:45-65) that the reference's own Python engine issues for a BASELINE.json config, together with what
the reference C++ simulator returned.  Replaying the calls on another backend must reproduce the returns: measurement
bits exactly (same rnd_seed), amplitudes / probabilities / expectation values within 1e-12.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

import numpy as np  # noqa: E402

from tests import refenv  # noqa: E402

refenv.import_projectq()

import projectq.backends._sim._cppsim as ref_cppsim  # noqa: E402
from projectq import MainEngine  # noqa: E402
from projectq.backends import Simulator  # noqa: E402
from projectq.cengines import AutoReplacer, DecompositionRuleSet, InstructionFilter, LocalOptimizer, TagRemover  # noqa: E402
from projectq.meta import Control  # noqa: E402
from projectq.ops import QFT, All, BasicMathGate, H, Measure, QubitOperator, R, Ry, Swap, TimeEvolution, X, get_inverse  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def enc(o):
    if isinstance(o, complex):
        return {"c": [o.real, o.imag]}
    if isinstance(o, (np.floating, np.integer)):
        return o.item()
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, (bool, int, float, str)) or o is None:
        return o
    if isinstance(o, (list, tuple)):
        return [enc(x) for x in o]
    raise TypeError(type(o))


class Recorder:
    """wraps the reference _cppsim.Simulator and logs every call with its return value"""

    RETURNS = {"measure_qubits", "get_probability", "get_amplitude", "get_expectation_value", "is_classical",
               "get_classical_value"}

    def __init__(self, seed, trace):
        self._inner = ref_cppsim.Simulator(seed)
        self._trace = trace

    def __getattr__(self, name):
        f = getattr(self._inner, name)

        def call(*args):
            if name == "emulate_math":
                raise RuntimeError("generic emulate_math is not recorded (python callable)")
            ret = f(*args)
            entry = {"m": name, "a": enc(list(args))}
            if name in self.RETURNS:
                entry["r"] = enc(ret)
            if name == "cheat":
                return ret
            self._trace.append(entry)
            return ret

        return call


def sample_state(sim_engine, trace, n_samples, rng):
    """record amplitudes at seeded sample indices (as a get_amplitudes-style check) + the norm"""
    mapping, vec = sim_engine._simulator._inner.cheat()
    vec = np.asarray(vec)
    idx = sorted(int(i) for i in rng.choice(len(vec), size=min(n_samples, len(vec)), replace=False))
    trace.append({"m": "_check_amplitudes", "a": [idx], "r": [[vec[i].real, vec[i].imag] for i in idx],
                  "map": {str(k): int(v) for k, v in dict(mapping).items()}})


def qft20():
    trace = []
    sim = Simulator(gate_fusion=True, rnd_seed=1)
    sim._simulator = Recorder(1, trace)
    eng = MainEngine(sim)  # default engine list
    q = eng.allocate_qureg(20)
    rng = np.random.default_rng(20)
    for i in range(20):  # seeded product-state preparation so that the QFT output has structure
        Ry(float(rng.uniform(0, np.pi))) | q[i]
    QFT | q
    eng.flush()
    sample_state(sim, trace, 256, rng)
    All(Measure) | q
    eng.flush()
    bits = [int(b) for b in q]
    return {"config": "QFT on 20 qubits + Measure, Simulator(gate_fusion=True, rnd_seed=1), default engine list, seeded Ry prep",
            "measured_bits": bits, "trace": trace}


def shor(N=4087, a=7, seed=3):
    """examples/shor.py run_shor with emulation on (the InstructionFilter lets BasicMathGate through)"""
    import projectq.libs.math
    import projectq.setups.decompositions
    from projectq.libs.math import AddConstant, AddConstantModN, MultiplyByConstantModN

    trace = []
    sim = Simulator(gate_fusion=True, rnd_seed=seed)
    sim._simulator = Recorder(seed, trace)

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
    # body of run_shor (examples/shor.py:31-86), restated
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
    xbits = [int(b) for b in x]
    return {"config": "Shor N=%d a=%d via examples/shor.py's run_shor with emulate_math on, rnd_seed=%d" % (N, a, seed),
            "measurements": measurements, "x_bits": xbits, "trace": trace}


def tfim(n=12, seed=4):
    trace = []
    sim = Simulator(gate_fusion=True, rnd_seed=seed)
    sim._simulator = Recorder(seed, trace)
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
    sample_state(sim, trace, 256, rng)
    p = sim.get_probability("010", q[:3])
    All(Measure) | q
    eng.flush()
    return {"config": "%d-qubit open-chain TFIM (J=1, h=0.7): Ry prep, 2 x [TimeEvolution(0.3, H), <H>]" % n,
            "energies": energies, "p010": p, "bits": [int(b) for b in q], "trace": trace}


def brickwork20(seed=6):
    """BASELINE config 2 at a size the reference finishes in seconds: same generator as bench.py (tests/helpers.py),
    20 qubits, depth 20, driven through the reference engine with gate_fusion=True and an empty engine list"""
    from projectq.ops import CNOT, CZ, Rx, Ry, Rz

    n, depth = 20, 20
    trace = []
    sim = Simulator(gate_fusion=True, rnd_seed=seed)
    sim._simulator = Recorder(seed, trace)
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
    sample_state(sim, trace, 512, rng)
    p = sim.get_probability("0110", q[3:7])
    All(Measure) | q
    eng.flush()
    return {"config": "random brickwork circuit, 20 qubits, depth 20 (Rx/Ry/Rz + CNOT/CZ), gate_fusion=True, rnd_seed=%d" % seed,
            "p0110": p, "bits": [int(b) for b in q], "trace": trace}


def main():
    for name, fn in (("qft20", qft20), ("shor4087", shor), ("tfim12", tfim), ("brickwork20", brickwork20)):
        data = fn()
        path = os.path.join(OUT, name + ".json")
        with open(path, "w") as f:
            json.dump(data, f, separators=(",", ":"))
        print(name, len(data["trace"]), "calls ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
