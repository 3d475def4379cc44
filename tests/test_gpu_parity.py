"""Parity of the CUDA engine (through the pybind11 shim over the C ABI) against the oracle.

Checker = the compiled, unmodified reference (oracle/_ref/_cppsim, when it loads) and the NumPy restatement
(oracle/statevec_oracle.py).  Tolerances from BASELINE.json: amplitudes 1e-12 max-abs, 1 - fidelity < 1e-12,
measurement outcomes identical for the same seed, emulate_math and qubit bookkeeping bit-exact.
"""
import math

import numpy as np
import pytest

from oracle.statevec_oracle import OracleSimulator
from tests.conftest import load_ref_cppsim
from tests.helpers import fidelity_gap, rand_state, rand_unitary, tfim_terms

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(scope="module")
def Backend():
    from projectq_b200.backend import SimulatorBackend

    return SimulatorBackend


def checkers():
    out = ["oracle"]
    if load_ref_cppsim() is not None:
        out.append("reference")
    return out


@pytest.fixture(params=checkers())
def make_checker(request):
    if request.param == "reference":
        mod = load_ref_cppsim()
        return lambda seed=1: mod.Simulator(seed)
    return lambda seed=1: OracleSimulator(seed)


def assert_same_state(gpu, chk, tol=TOL):
    m1, v1 = gpu.cheat()
    m2, v2 = chk.cheat()
    assert dict(m1) == dict(m2)
    v1 = np.asarray(v1)
    v2 = np.asarray(v2)
    assert v1.shape == v2.shape
    assert np.max(np.abs(v1 - v2)) < tol
    if np.vdot(v2, v2).real > 1e-20:
        assert fidelity_gap(v1, v2) < tol


def mat_arg(chk, m):
    return m if isinstance(chk, OracleSimulator) else m.tolist()


def wf_arg(chk, wf):
    return wf if isinstance(chk, OracleSimulator) else list(wf)


def prepare(Backend, make_checker, n, rng, seed=1, ids=None, scramble=True, **kw):
    gpu, chk = Backend(seed, **kw), make_checker(seed)
    ids = list(range(n)) if ids is None else ids
    for q in ids:
        gpu.allocate_qubit(q)
        chk.allocate_qubit(q)
    if scramble:
        wf = rand_state(rng, n)
        order = [int(x) for x in rng.permutation(ids)]
        gpu.set_wavefunction(wf, order)
        chk.set_wavefunction(wf_arg(chk, wf), order)
    return gpu, chk, ids


@pytest.mark.parametrize("fused", [False, True])
def test_random_circuits(Backend, make_checker, fused):
    rng = np.random.default_rng(101)
    for trial in range(8):
        n = int(rng.integers(6, 12))
        ids = [int(x) for x in rng.permutation(50)[:n]]
        gpu, chk, ids = prepare(Backend, make_checker, n, rng, ids=ids)
        for g in range(60):
            k = int(rng.integers(1, min(5, n - 1) + 1))
            nc = int(rng.integers(0, min(3, n - k) + 1))
            qs = [int(x) for x in rng.permutation(ids)[: k + nc]]
            m = rand_unitary(rng, k)
            gpu.apply_controlled_gate(m, qs[:k], qs[k:])
            chk.apply_controlled_gate(mat_arg(chk, m), qs[:k], qs[k:])
            if not fused:
                gpu.run()
                chk.run()
        assert_same_state(gpu, chk)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_dense_gate_every_placement_class(Backend, k):
    """low / high / spread target positions and control placements on an 18-qubit state, against the NumPy oracle"""
    rng = np.random.default_rng(200 + k)
    n = 18
    wf = rand_state(rng, n)
    placements = [list(range(k)), list(range(n - k, n)), sorted(int(x) for x in rng.permutation(n)[:k]),
                  [0] + list(range(n - k + 1, n))]
    # runs of low target bits shorter than k (the warp-transposing kernel with extra high targets)
    for c in range(2, k):
        placements.append(list(range(c)) + list(range(n - (k - c), n)))
        placements.append(list(range(c)) + list(range(c + 5, c + 5 + (k - c))))
    for targets in placements:
        free = [q for q in range(n) if q not in targets]
        for ctrls in ([], [free[0]], [free[-1], free[2]], [free[1], free[5], free[-2]], [free[-1]], [free[-3], free[-1]]):
            gpu, chk = Backend(1), OracleSimulator(1)
            for q in range(n):
                gpu.allocate_qubit(q)
                chk.allocate_qubit(q)
            gpu.set_wavefunction(wf, list(range(n)))
            chk.set_wavefunction(wf, list(range(n)))
            m = rand_unitary(rng, k)
            tq = [int(x) for x in rng.permutation(targets)]  # caller order != position order
            gpu.apply_controlled_gate(m, tq, ctrls)
            gpu.run()
            chk.apply_controlled_gate(m, tq, ctrls)
            assert_same_state(gpu, chk)


def test_diagonal_passes(Backend, make_checker):
    rng = np.random.default_rng(7)
    gpu, chk, ids = prepare(Backend, make_checker, 10, rng)
    for g in range(40):
        k = int(rng.integers(1, 4))
        qs = [int(x) for x in rng.permutation(ids)[: k + 1]]
        d = np.diag(np.exp(1j * rng.uniform(0, 2 * np.pi, 1 << k)))
        gpu.apply_controlled_gate(d, qs[:k], qs[k:])
        chk.apply_controlled_gate(mat_arg(chk, d), qs[:k], qs[k:])
        if g % 7 == 0:
            gpu.run()
    assert gpu.stats()["diag_passes"] > 0
    assert_same_state(gpu, chk)


def test_measure_collapse_deallocate(Backend, make_checker):
    rng = np.random.default_rng(11)
    for seed in (1, 2, 12345, 4294967295):
        gpu, chk, ids = prepare(Backend, make_checker, 9, rng, seed=seed)
        for g in range(20):
            k = int(rng.integers(1, 3))
            qs = [int(x) for x in rng.permutation(ids)[: k + 1]]
            m = rand_unitary(rng, k)
            gpu.apply_controlled_gate(m, qs[:k], qs[k:])
            chk.apply_controlled_gate(mat_arg(chk, m), qs[:k], qs[k:])
        for mids in ([3], [0, 5], [7, 1, 2]):
            assert list(gpu.measure_qubits(mids)) == list(chk.measure_qubits(mids))
            assert_same_state(gpu, chk)
        for q in (3, 0, 5, 2):
            assert gpu.is_classical(q, 1e-12) == chk.is_classical(q, 1e-12)
            assert gpu.get_classical_value(q, 1e-12) == chk.get_classical_value(q, 1e-12)
            gpu.deallocate_qubit(q)
            chk.deallocate_qubit(q)
            assert_same_state(gpu, chk)
        assert not gpu.is_classical(4, 1e-12)
        with pytest.raises(RuntimeError):
            gpu.deallocate_qubit(4)
        gpu.collapse_wavefunction([4, 6], [True, False])
        chk.collapse_wavefunction([4, 6], [True, False])
        assert_same_state(gpu, chk)
        # allocate again after deallocation: new qubit is the new top bit
        gpu.allocate_qubit(77)
        chk.allocate_qubit(77)
        m = rand_unitary(rng, 2)
        gpu.apply_controlled_gate(m, [77, 4], [])
        chk.apply_controlled_gate(mat_arg(chk, m), [77, 4], [])
        assert_same_state(gpu, chk)


def test_measurement_stream_many_draws(Backend, make_checker):
    """the host RNG stream is replayed draw by draw: 12 single-qubit measurements, re-superposed in between"""
    H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / math.sqrt(2)
    gpu, chk = Backend(7), make_checker(7)
    for q in range(12):
        gpu.allocate_qubit(q)
        chk.allocate_qubit(q)
    got, want = [], []
    for rounds in range(3):
        for q in range(12):
            gpu.apply_controlled_gate(H, [q], [])
            chk.apply_controlled_gate(mat_arg(chk, H), [q], [])
        for q in range(0, 12, 3):
            got += list(gpu.measure_qubits([q, (q + 1) % 12]))
            want += list(chk.measure_qubits([q, (q + 1) % 12]))
    assert got == want
    assert_same_state(gpu, chk)


def test_probability_amplitude(Backend, make_checker):
    rng = np.random.default_rng(13)
    gpu, chk, ids = prepare(Backend, make_checker, 11, rng)
    for _ in range(10):
        k = int(rng.integers(1, 6))
        sub = [int(x) for x in rng.permutation(ids)[:k]]
        bits = [bool(b) for b in rng.integers(0, 2, k)]
        assert abs(gpu.get_probability(bits, sub) - chk.get_probability(bits, sub)) < TOL
        full = [int(x) for x in rng.permutation(ids)]
        fb = [bool(b) for b in rng.integers(0, 2, len(ids))]
        assert abs(gpu.get_amplitude(fb, full) - chk.get_amplitude(fb, full)) < TOL
    with pytest.raises(RuntimeError):
        gpu.get_probability([True], [999])
    with pytest.raises(RuntimeError):
        gpu.get_amplitude([True] * 3, ids[:3])


def test_pauli_operators(Backend, make_checker):
    rng = np.random.default_rng(3)
    n = 9
    gpu, chk, ids = prepare(Backend, make_checker, n, rng)
    order = [int(x) for x in rng.permutation(ids)]
    terms = tfim_terms(n) + [([(0, "Y"), (3, "X"), (5, "Z")], 0.37), ([], 0.25), ([(2, "Y"), (6, "Y")], -1.1),
                             ([(1, "X"), (2, "Y"), (4, "Z"), (7, "Y"), (8, "X")], 0.5)]
    e1 = gpu.get_expectation_value(terms, order)
    e2 = chk.get_expectation_value(terms, order)
    assert abs(e1 - e2) < TOL * sum(abs(c) for _, c in terms)
    cterms = [(t, c * (1 + 0.5j)) for t, c in terms]
    gpu.apply_qubit_operator(cterms, order)
    chk.apply_qubit_operator(cterms, order)
    assert_same_state(gpu, chk, tol=1e-11)  # unnormalised result, norm ~ ||H|| ~ 20
    with pytest.raises(TypeError):
        gpu.get_expectation_value([([(0, "X")], 1j)], order)


def test_pauli_operators_tiled_paths(Backend, make_checker):
    """a state wider than one shared-memory tile (16 qubits > 11 tile bits): X/Y terms on every bit (several tile-bit sets),
    more than 64 terms in one launch group, and terms whose X support is too wide for any tile (global gathers)"""
    rng = np.random.default_rng(33)
    n = 16
    gpu, chk, ids = prepare(Backend, make_checker, n, rng)
    order = [int(x) for x in rng.permutation(ids)]
    terms = tfim_terms(n)
    for _ in range(70):  # random strings of weight 1..5
        w = int(rng.integers(1, 6))
        qs = sorted(int(x) for x in rng.permutation(n)[:w])
        terms.append(([(q, "XYZ"[int(rng.integers(0, 3))]) for q in qs], float(rng.normal())))
    terms.append(([(q, "X") for q in range(13)], 0.3))                       # wider than a tile
    terms.append(([(q, "Y" if q % 2 else "X") for q in range(2, 16)], -0.2))  # 14 X/Y factors
    terms.append(([], 0.4))
    h1 = sum(abs(c) for _, c in terms)
    e1 = gpu.get_expectation_value(terms, order)
    e2 = chk.get_expectation_value(terms, order)
    assert abs(e1 - e2) < TOL * h1
    sub = terms[:40] + terms[-3:]
    gpu.emulate_time_evolution(sub, 0.05, order, [])
    chk.emulate_time_evolution(sub, 0.05, order, [])
    assert_same_state(gpu, chk)
    cterms = [(t, c * (0.3 - 0.2j)) for t, c in terms]
    gpu.apply_qubit_operator(cterms, order)
    chk.apply_qubit_operator(cterms, order)
    assert_same_state(gpu, chk, tol=1e-11 * h1)


@pytest.mark.parametrize("ctrl", [[], [8]])
def test_time_evolution(Backend, make_checker, ctrl):
    rng = np.random.default_rng(5)
    n = 9
    gpu, chk, ids = prepare(Backend, make_checker, n, rng)
    terms = tfim_terms(n - 1) + [([], 0.3), ([(0, "Y"), (4, "Y")], 0.2)]
    qs = [q for q in ids if q != 8]
    gpu.emulate_time_evolution(terms, 0.8, qs, ctrl)
    chk.emulate_time_evolution(terms, 0.8, qs, ctrl)
    assert_same_state(gpu, chk)
    with pytest.raises(TypeError):
        gpu.emulate_time_evolution([([(0, "X")], 1j)], 0.1, qs, [])


def test_time_evolution_matches_scipy_expm(Backend):
    """independent numerical check like the reference's own test (_simulator_test.py:514-570)"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    rng = np.random.default_rng(17)
    n = 8
    wf = rand_state(rng, n)
    gpu = Backend(1)
    for q in range(n):
        gpu.allocate_qubit(q)
    gpu.set_wavefunction(wf, list(range(n)))
    terms = tfim_terms(n, J=0.8, h=1.3) + [([], 0.5)]
    paulis = {"X": sp.csr_matrix([[0, 1], [1, 0]]), "Y": sp.csr_matrix([[0, -1j], [1j, 0]]),
              "Z": sp.csr_matrix([[1, 0], [0, -1]])}
    Hm = sp.csr_matrix((1 << n, 1 << n), dtype=np.complex128)
    for term, c in terms:
        op = sp.identity(1, dtype=np.complex128, format="csr")
        tdict = dict(term)
        for q in range(n - 1, -1, -1):  # qubit 0 = least-significant bit
            op = sp.kron(op, paulis[tdict[q]] if q in tdict else sp.identity(2, format="csr"), format="csr")
        Hm = Hm + c * op
    t = 1.7
    want = spl.expm_multiply(-1j * t * Hm, wf)
    gpu.emulate_time_evolution(terms, t, list(range(n)), [])
    got = np.asarray(gpu.cheat()[1])
    assert np.max(np.abs(got - want)) < 1e-10  # scipy's own truncation error dominates


def test_emulate_math_bit_exact(Backend, make_checker):
    n = 8
    for ctrl in ([], [7], [6, 7]):
        for which in range(6):
            gpu, chk = Backend(1), make_checker(1)
            for q in range(n):
                gpu.allocate_qubit(q)
                chk.allocate_qubit(q)
            wf = np.arange(1, (1 << n) + 1, dtype=np.float64) + 1j * np.arange((1 << n), 0, -1)
            order = [2, 0, 1, 3, 5, 4, 6, 7]
            gpu.set_wavefunction(wf, order)
            chk.set_wavefunction(wf_arg(chk, wf), order)
            regs = [[0, 1, 2], [3, 4, 5]] if which >= 4 else [[0, 1, 2, 3, 4]]
            for s in (gpu, chk):
                if which == 0:
                    s.emulate_math_addConstant(5, regs, ctrl)
                elif which == 1:
                    s.emulate_math_addConstant(-3, regs, ctrl)
                elif which == 2:
                    s.emulate_math_addConstantModN(4, 32, regs, ctrl)
                elif which == 3:
                    s.emulate_math_multiplyByConstantModN(7, 32, regs, ctrl)
                elif which == 4:
                    s.emulate_math(lambda x: [x[0], (x[1] + x[0]) % 8], regs, ctrl)
                else:
                    s.emulate_math(lambda x: [(x[0] - 3), (x[1] * 3) % 8], regs, ctrl)
            m1, v1 = gpu.cheat()
            m2, v2 = chk.cheat()
            assert dict(m1) == dict(m2)
            assert np.array_equal(np.asarray(v1), np.asarray(v2))  # bit-exact


def test_reference_math_kats(Backend):
    """known answers of the reference's own test (_simulator_test.py:739-786): 5-qubit register starting at 0"""
    want = {"add3": [1, 1, 0, 0, 0], "add4mod5": [0, 1, 0, 0, 0], "mul15mod16": [0, 1, 1, 1, 0]}
    X = np.array([[0, 1], [1, 0]], dtype=np.complex128)

    def measured(sim):
        return [int(b) for b in sim.measure_qubits(list(range(5)))]

    s = Backend(1)
    for q in range(5):
        s.allocate_qubit(q)
    s.emulate_math_addConstant(3, [list(range(5))], [])
    assert measured(s) == want["add3"]
    # qureg now holds 3; (3 + 4) % 5 = 2
    s.emulate_math_addConstantModN(4, 5, [list(range(5))], [])
    assert measured(s) == want["add4mod5"]
    # 2 * 15 % 16 = 14
    s.emulate_math_multiplyByConstantModN(15, 16, [list(range(5))], [])
    assert measured(s) == want["mul15mod16"]
    del X


def test_error_classes(Backend):
    s = Backend(1)
    s.allocate_qubit(0)
    with pytest.raises(RuntimeError):
        s.allocate_qubit(0)
    with pytest.raises(ValueError):
        s.collapse_wavefunction([0], [True, False])
    with pytest.raises(RuntimeError):
        s.collapse_wavefunction([5], [True])
    with pytest.raises(RuntimeError):
        s.collapse_wavefunction([0], [True])  # probability 0
    with pytest.raises(RuntimeError):
        s.set_wavefunction(np.ones(4) / 2, [0, 1])
    with pytest.raises(ValueError):
        s.apply_controlled_gate(np.eye(64), [0, 1, 2, 3, 4, 5], [])
    # a failing call must not poison the queue (the reference does, simulator.hpp:522-526)
    s.apply_controlled_gate(np.array([[0, 1], [1, 0]], dtype=complex), [0], [])
    assert abs(s.get_probability([True], [0]) - 1.0) < TOL


def test_state_growth_and_shrink(Backend, make_checker):
    """allocate up to 16 qubits one by one with gates in between, then measure and deallocate in mixed order"""
    rng = np.random.default_rng(23)
    gpu, chk = Backend(3), make_checker(3)
    n = 16
    for q in range(n):
        gpu.allocate_qubit(q)
        chk.allocate_qubit(q)
        m = rand_unitary(rng, 1)
        gpu.apply_controlled_gate(m, [q], [q - 1] if q else [])
        chk.apply_controlled_gate(mat_arg(chk, m), [q], [q - 1] if q else [])
    assert_same_state(gpu, chk)
    order = [int(x) for x in rng.permutation(n)]
    for q in order[:10]:
        assert list(gpu.measure_qubits([q])) == list(chk.measure_qubits([q]))
        gpu.deallocate_qubit(q)
        chk.deallocate_qubit(q)
    assert_same_state(gpu, chk)


def test_wide_state_sampled_amplitudes_vs_numpy(Backend):
    """22 qubits: fused brickwork circuit against the NumPy oracle on the full state"""
    from tests.helpers import brickwork_circuit, pack_gate_stream

    n = 22
    gates = brickwork_circuit(n, 4, seed=5)
    gpu, chk = Backend(1), OracleSimulator(1)
    for q in range(n):
        gpu.allocate_qubit(q)
        chk.allocate_qubit(q)
    body, cnt = pack_gate_stream(gates)
    gpu.apply_gate_stream(body, cnt, True)
    for m, t, c in gates:
        chk.apply_controlled_gate(m, t, c)
    assert_same_state(gpu, chk)
    st = gpu.stats()
    assert st["gates_ingested"] == len(gates)
    assert sum(st["dense_passes"]) + st["diag_passes"] < len(gates) / 3  # the fuser really fuses
