"""Pins the NumPy oracle (oracle/statevec_oracle.py) to the compiled, unmodified reference
(oracle/_ref/_cppsim, built from /root/reference by oracle/Makefile) and to libstdc++'s RNG stream."""
import math

import numpy as np
import pytest

from oracle.statevec_oracle import MT19937, OracleSimulator

TOL = 1e-12


def rand_unitary(rng, k):
    d = 1 << k
    a = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def both(ref_cppsim, seed=1):
    return ref_cppsim.Simulator(seed), OracleSimulator(seed)


def state_close(ref, orc):
    m1, v1 = ref.cheat()
    m2, v2 = orc.cheat()
    assert dict(m1) == dict(m2)
    v1 = np.asarray(v1)
    assert v1.shape == v2.shape
    assert np.max(np.abs(v1 - v2)) < TOL


def test_rng_golden_stream():
    # SURVEY §7 hard part 1: libstdc++ stream for seed 1
    r = MT19937(1)
    got = [r.uniform01() for _ in range(4)]
    want = [0.99718480823026556, 0.93255736136816547, 0.128124447772306, 0.99904051546527362]
    assert got == want


@pytest.mark.parametrize("fused", [False, True])
def test_random_circuits_vs_reference(ref_cppsim, fused):
    rng = np.random.default_rng(7)
    for trial in range(6):
        ref, orc = both(ref_cppsim)
        n = 9
        ids = list(rng.permutation(40)[:n])
        ids = [int(x) for x in ids]
        for q in ids:
            ref.allocate_qubit(q)
            orc.allocate_qubit(q)
        # scramble the id -> position map
        wf = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        wf /= np.linalg.norm(wf)
        order = [int(x) for x in rng.permutation(ids)]
        ref.set_wavefunction(list(wf), order)
        orc.set_wavefunction(wf, order)
        for g in range(40):
            k = int(rng.integers(1, 4))
            nc = int(rng.integers(0, 3))
            qs = [int(x) for x in rng.permutation(ids)[: k + nc]]
            m = rand_unitary(rng, k)
            ref.apply_controlled_gate(m.tolist(), qs[:k], qs[k:])
            if not fused:
                ref.run()
            orc.apply_controlled_gate(m, qs[:k], qs[k:])
        state_close(ref, orc)
        # probabilities / amplitudes
        sub = [int(x) for x in rng.permutation(ids)[:3]]
        bits = [bool(b) for b in rng.integers(0, 2, 3)]
        assert abs(ref.get_probability(bits, sub) - orc.get_probability(bits, sub)) < TOL
        full = [int(x) for x in rng.permutation(ids)]
        fb = [bool(b) for b in rng.integers(0, 2, n)]
        assert abs(ref.get_amplitude(fb, full) - orc.get_amplitude(fb, full)) < TOL


def test_measure_collapse_dealloc_vs_reference(ref_cppsim):
    rng = np.random.default_rng(11)
    for seed in (1, 2, 12345):
        ref, orc = both(ref_cppsim, seed)
        n = 8
        for q in range(n):
            ref.allocate_qubit(q)
            orc.allocate_qubit(q)
        for q in range(n):  # make sure every qubit leaves |0>
            m = rand_unitary(rng, 1)
            ref.apply_controlled_gate(m.tolist(), [q], [])
            orc.apply_controlled_gate(m, [q], [])
        for g in range(30):
            k = int(rng.integers(1, 3))
            qs = [int(x) for x in rng.permutation(n)[: k + 1]]
            m = rand_unitary(rng, k)
            ref.apply_controlled_gate(m.tolist(), qs[:k], qs[k:])
            orc.apply_controlled_gate(m, qs[:k], qs[k:])
        for ids in ([3], [0, 5], [7, 1, 2]):
            assert list(ref.measure_qubits(ids)) == orc.measure_qubits(ids)
            state_close(ref, orc)
        for q in (3, 0, 5):
            assert ref.is_classical(q, 1e-12) == orc.is_classical(q)
            assert ref.get_classical_value(q, 1e-12) == orc.get_classical_value(q)
            ref.deallocate_qubit(q)
            orc.deallocate_qubit(q)
            state_close(ref, orc)
        with pytest.raises(RuntimeError):
            orc.deallocate_qubit(4)
        with pytest.raises(RuntimeError):
            ref.deallocate_qubit(4)
        ref.collapse_wavefunction([4, 6], [True, False])
        orc.collapse_wavefunction([4, 6], [True, False])
        state_close(ref, orc)


def tfim_terms(n, J=1.0, h=0.7):
    terms = [([(i, "Z"), (i + 1, "Z")], -J) for i in range(n - 1)]
    terms += [([(i, "X")], -h) for i in range(n)]
    return terms


def test_pauli_ops_vs_reference(ref_cppsim):
    rng = np.random.default_rng(3)
    n = 7
    ref, orc = both(ref_cppsim)
    for q in range(n):
        ref.allocate_qubit(q)
        orc.allocate_qubit(q)
    wf = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    wf /= np.linalg.norm(wf)
    order = [int(x) for x in rng.permutation(n)]
    ref.set_wavefunction(list(wf), order)
    orc.set_wavefunction(wf, order)
    ids = [int(x) for x in rng.permutation(n)]
    terms = tfim_terms(n) + [([(0, "Y"), (3, "X"), (5, "Z")], 0.37), ([], 0.25), ([(2, "Y"), (6, "Y")], -1.1)]
    e1 = ref.get_expectation_value(terms, ids)
    e2 = orc.get_expectation_value(terms, ids)
    assert abs(e1 - e2) < TOL
    cterms = [(t, c * (1 + 0.5j)) for t, c in terms]
    ref.apply_qubit_operator(cterms, ids)
    orc.apply_qubit_operator(cterms, ids)
    state_close(ref, orc)


@pytest.mark.parametrize("ctrl", [[], [7]])
def test_time_evolution_vs_reference(ref_cppsim, ctrl):
    rng = np.random.default_rng(5)
    n = 8
    ref, orc = both(ref_cppsim)
    for q in range(n):
        ref.allocate_qubit(q)
        orc.allocate_qubit(q)
    wf = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    wf /= np.linalg.norm(wf)
    ref.set_wavefunction(list(wf), list(range(n)))
    orc.set_wavefunction(wf, list(range(n)))
    terms = tfim_terms(n - 1) + [([], 0.3)]
    ids = list(range(n - 1))
    ref.emulate_time_evolution(terms, 0.8, ids, ctrl)
    orc.emulate_time_evolution(terms, 0.8, ids, ctrl)
    state_close(ref, orc)


def test_emulate_math_vs_reference(ref_cppsim):
    n = 7
    for ctrl in ([], [6]):
        for which in range(5):
            ref, orc = both(ref_cppsim)
            for q in range(n):
                ref.allocate_qubit(q)
                orc.allocate_qubit(q)
            wf = np.arange(1, (1 << n) + 1, dtype=np.float64) + 0j  # distinct, exactly representable
            order = [2, 0, 1, 3, 5, 4, 6]
            ref.set_wavefunction(list(wf), order)
            orc.set_wavefunction(wf, order)
            regs = [[0, 1, 2], [3, 4, 5]] if which == 4 else [[0, 1, 2, 3, 4]]
            for s in (ref, orc):
                if which == 0:
                    s.emulate_math_addConstant(5, regs, ctrl)
                elif which == 1:
                    s.emulate_math_addConstant(-3, regs, ctrl)
                elif which == 2:
                    s.emulate_math_addConstantModN(4, 29, regs, ctrl)
                elif which == 3:
                    s.emulate_math_multiplyByConstantModN(7, 31, regs, ctrl)
                else:
                    s.emulate_math(lambda x: [x[0], (x[1] + x[0]) % 8], regs, ctrl)
            m1, v1 = ref.cheat()
            m2, v2 = orc.cheat()
            assert dict(m1) == m2
            assert np.array_equal(np.asarray(v1), v2)  # bit-exact
