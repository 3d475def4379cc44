"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py [--out DIR]

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
from projectq.backends import Simulator  # noqa: E402

from tests.golden import programs  # noqa: E402

OUT = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else os.path.dirname(os.path.abspath(__file__))


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
    idx = programs.sample_indices(len(vec), n_samples, rng)
    trace.append({"m": "_check_amplitudes", "a": [idx], "r": [[vec[i].real, vec[i].imag] for i in idx],
                  "map": {str(k): int(v) for k, v in dict(mapping).items()}})


CONFIGS = {
    "qft20": "QFT on 20 qubits + Measure, Simulator(gate_fusion=True, rnd_seed=1), default engine list, seeded Ry prep",
    "shor4087": "Shor N=4087 a=7 via examples/shor.py's run_shor with emulate_math on, rnd_seed=3",
    "tfim12": "12-qubit open-chain TFIM (J=1, h=0.7): Ry prep, 2 x [TimeEvolution(0.3, H), <H>]",
    "brickwork20": "random brickwork circuit, 20 qubits, depth 20 (Rx/Ry/Rz + CNOT/CZ), gate_fusion=True, rnd_seed=6",
}


def record(name):
    """run one program of tests/golden/programs.py on the reference engine + reference C++ simulator, recording the seam"""
    trace = []

    def make_sim(gate_fusion, rnd_seed):
        sim = Simulator(gate_fusion=gate_fusion, rnd_seed=rnd_seed)
        sim._simulator = Recorder(rnd_seed, trace)
        return sim

    out = {"config": CONFIGS[name]}
    for step in programs.PROGRAMS[name](make_sim):
        if step[0] == "state":
            _, sim, _, rng = step
            sample_state(sim, trace, programs.SAMPLES[name], rng)
        else:
            out.update(step[1])
    out["trace"] = trace
    return out


def main():
    for name in programs.PROGRAMS:
        data = record(name)
        path = os.path.join(OUT, name + ".json")
        with open(path, "w") as f:
            json.dump(data, f, separators=(",", ":"))
        print(name, len(data["trace"]), "calls ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
