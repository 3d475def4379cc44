"""Replays the golden call traces recorded from the unmodified reference (tests/golden/*.json, generator
tests/golden/make_golden.py) on (a) the NumPy oracle — CPU, pins the oracle — and (b) the CUDA engine — GPU."""
import json
import os

import numpy as np
import pytest

from oracle.statevec_oracle import OracleSimulator

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-12


def load(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def dec(o):
    if isinstance(o, dict) and "c" in o:
        return complex(o["c"][0], o["c"][1])
    if isinstance(o, list):
        return [dec(x) for x in o]
    return o


def replay(sim, trace, amplitudes):
    n_checked = 0
    for e in trace:
        m, args = e["m"], dec(e["a"])
        if m == "_check_amplitudes":
            idx = args[0]
            mapping, got = amplitudes(sim, idx)
            assert {str(k): int(v) for k, v in dict(mapping).items()} == e["map"]
            want = np.array([complex(r, i) for r, i in e["r"]])
            assert np.max(np.abs(np.asarray(got) - want)) < TOL
            n_checked += len(idx)
            continue
        if m == "apply_controlled_gate":
            args[0] = np.array(args[0], dtype=np.complex128)
        ret = getattr(sim, m)(*args)
        if "r" in e:
            want = dec(e["r"])
            if m == "measure_qubits":
                assert [bool(b) for b in ret] == [bool(b) for b in want], (m, args)
            elif isinstance(want, bool):
                assert bool(ret) == want
            else:
                assert abs(complex(ret) - complex(want)) < TOL * max(1.0, abs(complex(want))), (m, ret, want)
            n_checked += 1
    return n_checked


def oracle_amplitudes(sim, idx):
    mapping, vec = sim.cheat()
    return mapping, vec[idx]


def gpu_amplitudes(sim, idx):
    mapping, _ = sim.cheat() if sim.num_qubits() <= 22 else ({}, None)
    return mapping, sim.get_amplitudes(np.array(idx, dtype=np.uint64))


@pytest.mark.parametrize("name", ["shor4087", "tfim12"])
def test_oracle_reproduces_reference_golden(name):
    data = load(name)
    assert replay(OracleSimulator(_seed(data)), data["trace"], oracle_amplitudes) > 10


def test_oracle_reproduces_qft20_golden():
    data = load("qft20")  # 2^20 amplitudes through NumPy gathers
    assert replay(OracleSimulator(1), data["trace"], oracle_amplitudes) > 10


def test_oracle_reproduces_brickwork20_golden():
    data = load("brickwork20")  # 590 gates on 2^20 amplitudes
    assert replay(OracleSimulator(_seed(data)), data["trace"], oracle_amplitudes) > 10


def _seed(data):
    import re

    m = re.search(r"rnd_seed=(\d+)", data["config"])
    return int(m.group(1)) if m else 4


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["qft20", "shor4087", "tfim12", "brickwork20"])
@pytest.mark.parametrize("fusion", [0, 1, 5])
def test_cuda_engine_reproduces_reference_golden(name, fusion):
    from projectq_b200.backend import SimulatorBackend

    data = load(name)
    sim = SimulatorBackend(_seed(data), fusion_max_qubits=fusion)
    assert replay(sim, data["trace"], gpu_amplitudes) > 10
