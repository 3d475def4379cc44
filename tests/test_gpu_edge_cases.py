"""Edge cases of the native seam on the GPU, against the compiled reference / the NumPy oracle: empty and tiny states, empty
argument lists, RNG consumption, non-classical probes, precondition-violating math gates, queue limits."""
import math

import numpy as np
import pytest

from oracle.statevec_oracle import OracleSimulator
from tests.conftest import load_ref_cppsim
from tests.helpers import rand_state, rand_unitary

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def Backend():
    from projectq_b200.backend import SimulatorBackend

    return SimulatorBackend


def checker(seed=1):
    mod = load_ref_cppsim()
    return mod.Simulator(seed) if mod is not None else OracleSimulator(seed)


def is_ref(chk):
    return not isinstance(chk, OracleSimulator)


def marg(chk, m):
    return m.tolist() if is_ref(chk) else m


def warg(chk, w):
    return list(w) if is_ref(chk) else w


def same(gpu, chk, tol=TOL):
    m1, v1 = gpu.cheat()
    m2, v2 = chk.cheat()
    assert dict(m1) == dict(m2)
    assert np.max(np.abs(np.asarray(v1) - np.asarray(v2))) < tol


def test_zero_qubit_state(Backend):
    s = Backend(1)
    mapping, vec = s.cheat()
    assert dict(mapping) == {} and list(vec) == [1 + 0j]
    assert s.measure_qubits([]) == []
    assert s.get_probability([], []) == 1.0
    assert s.get_amplitude([], []) == 1 + 0j
    assert s.get_expectation_value([([], 2.5)], []) == 2.5
    s.run()
    s.collapse_wavefunction([], [])
    s.set_wavefunction(np.array([1j]), [])
    assert s.get_amplitude([], []) == 1j


def test_single_qubit_everything(Backend):
    gpu, chk = Backend(9), checker(9)
    for s in (gpu, chk):
        s.allocate_qubit(42)
    h = np.array([[1, 1], [1, -1]], dtype=np.complex128) / math.sqrt(2)
    gpu.apply_controlled_gate(h, [42], [])
    chk.apply_controlled_gate(marg(chk, h), [42], [])
    same(gpu, chk)
    assert abs(gpu.get_probability([True], [42]) - 0.5) < TOL
    assert gpu.is_classical(42, 1e-12) is False
    assert gpu.get_classical_value(42, 1e-12) == chk.get_classical_value(42, 1e-12)
    assert list(gpu.measure_qubits([42])) == list(chk.measure_qubits([42]))
    same(gpu, chk)
    gpu.deallocate_qubit(42)
    chk.deallocate_qubit(42)
    same(gpu, chk)


def test_empty_measurement_consumes_one_draw(Backend):
    """the reference draws from its RNG once per measure_qubits call, even for an empty id list (simulator.hpp:153)"""
    gpu, chk = Backend(5), checker(5)
    h = np.array([[1, 1], [1, -1]], dtype=np.complex128) / math.sqrt(2)
    for q in range(6):
        gpu.allocate_qubit(q)
        chk.allocate_qubit(q)
        gpu.apply_controlled_gate(h, [q], [])
        chk.apply_controlled_gate(marg(chk, h), [q], [])
    assert list(gpu.measure_qubits([])) == list(chk.measure_qubits([])) == []
    same(gpu, chk)
    for q in range(6):
        assert list(gpu.measure_qubits([q])) == list(chk.measure_qubits([q]))
    # duplicate ids in one call
    for q in range(6):
        gpu.apply_controlled_gate(h, [q], [])
        chk.apply_controlled_gate(marg(chk, h), [q], [])
    assert list(gpu.measure_qubits([2, 2, 4])) == list(chk.measure_qubits([2, 2, 4]))
    same(gpu, chk)


def test_measurement_of_unnormalised_state_takes_last_index(Backend):
    """if the running sum never reaches the draw the reference ends on the last index (simulator.hpp:156-160);
    seed 1 draws 0.997... first"""
    gpu, chk = Backend(1), checker(1)
    n = 5
    for q in range(n):
        gpu.allocate_qubit(q)
        chk.allocate_qubit(q)
    wf = np.full(1 << n, 0.1 + 0j)  # total probability 0.32
    gpu.set_wavefunction(wf, list(range(n)))
    chk.set_wavefunction(warg(chk, wf), list(range(n)))
    a, b = list(gpu.measure_qubits([0, 3, 4])), list(chk.measure_qubits([0, 3, 4]))
    assert a == b == [True, True, True]
    same(gpu, chk)


def test_get_classical_value_on_superpositions(Backend):
    rng = np.random.default_rng(8)
    n = 7
    for trial in range(6):
        gpu, chk = Backend(1), checker(1)
        for q in range(n):
            gpu.allocate_qubit(q)
            chk.allocate_qubit(q)
        wf = rand_state(rng, n)
        wf[rng.random(1 << n) < 0.6] = 0  # many exact zeros so that the scan order matters
        wf /= np.linalg.norm(wf)
        order = [int(x) for x in rng.permutation(n)]
        gpu.set_wavefunction(wf, order)
        chk.set_wavefunction(warg(chk, wf), order)
        gpu.deallocate_qubit  # noqa: B018  (no call: just keep both objects alive)
        for q in range(n):
            assert gpu.is_classical(q, 1e-12) == chk.is_classical(q, 1e-12)
            assert gpu.get_classical_value(q, 1e-12) == chk.get_classical_value(q, 1e-12), (trial, q)


def test_operator_corner_cases(Backend):
    rng = np.random.default_rng(2)
    n = 6
    gpu, chk = Backend(1), checker(1)
    for q in range(n):
        gpu.allocate_qubit(q)
        chk.allocate_qubit(q)
    wf = rand_state(rng, n)
    gpu.set_wavefunction(wf, list(range(n)))
    chk.set_wavefunction(warg(chk, wf), list(range(n)))
    ids = list(range(n))
    assert gpu.get_expectation_value([], ids) == 0.0
    # identity-only Hamiltonian: a pure phase (simulator.hpp:396-400,431-436)
    gpu.emulate_time_evolution([([], 0.7)], 1.3, ids, [])
    chk.emulate_time_evolution([([], 0.7)], 1.3, ids, [])
    same(gpu, chk)
    gpu.emulate_time_evolution([], 0.5, ids, [2])
    chk.emulate_time_evolution([], 0.5, ids, [2])
    same(gpu, chk)
    # negative time, control, a 5-body term
    terms = [([(0, "X"), (1, "Y"), (2, "Z"), (3, "X"), (4, "Y")], 0.8), ([(5, "Z")], -0.3), ([], 0.1)]
    gpu.emulate_time_evolution(terms, -0.6, ids, [])
    chk.emulate_time_evolution(terms, -0.6, ids, [])
    same(gpu, chk)
    # empty operator annihilates the state
    gpu.apply_qubit_operator([], ids)
    chk.apply_qubit_operator([], ids)
    same(gpu, chk)
    assert np.all(np.asarray(gpu.cheat()[1]) == 0)


def test_math_gate_precondition_violations_accumulate(Backend):
    """(x + a) % N with register values >= N: several sources land on one target and are summed (simulator.hpp:261)"""
    n = 6
    gpu, chk = Backend(1), checker(1)
    for q in range(n):
        gpu.allocate_qubit(q)
        chk.allocate_qubit(q)
    wf = (np.arange(1, (1 << n) + 1) + 0j) / 100.0
    gpu.set_wavefunction(wf, list(range(n)))
    chk.set_wavefunction(warg(chk, wf), list(range(n)))
    for s in (gpu, chk):
        s.emulate_math_addConstantModN(3, 5, [[0, 1, 2, 3]], [5])
    same(gpu, chk, tol=1e-14)
    for s in (gpu, chk):
        s.emulate_math_multiplyByConstantModN(2, 6, [[0, 1, 2]], [])
    same(gpu, chk, tol=1e-14)
    # empty register list and an empty register are no-ops
    for s in (gpu, chk):
        s.emulate_math_addConstant(7, [], [])
        s.emulate_math_addConstant(7, [[]], [0])
    same(gpu, chk, tol=1e-14)


def test_queue_limit_and_many_controls(Backend):
    rng = np.random.default_rng(12)
    n = 12
    gpu, chk = Backend(1), OracleSimulator(1)
    for q in range(n):
        gpu.allocate_qubit(q)
        chk.allocate_qubit(q)
    h = np.array([[1, 1], [1, -1]], dtype=np.complex128) / math.sqrt(2)
    for q in range(n):
        gpu.apply_controlled_gate(h, [q], [])
        chk.apply_controlled_gate(h, [q], [])
    # more gates than the engine buffers before draining on its own (4096)
    gates = []
    for g in range(4500):
        q = int(rng.integers(0, n))
        th = float(rng.uniform(0, 6.28))
        m = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]], dtype=np.complex128)
        gates.append((m, q))
    for m, q in gates:
        gpu.apply_controlled_gate(m, [q], [])
        chk.apply_controlled_gate(m, [q], [])
    # a 10-fold controlled gate and a gate whose controls are all the other qubits
    m = rand_unitary(rng, 1)
    gpu.apply_controlled_gate(m, [11], list(range(10)))
    chk.apply_controlled_gate(m, [11], list(range(10)))
    m2 = rand_unitary(rng, 2)
    gpu.apply_controlled_gate(m2, [3, 7], [q for q in range(n) if q not in (3, 7)])
    chk.apply_controlled_gate(m2, [3, 7], [q for q in range(n) if q not in (3, 7)])
    same(gpu, chk, tol=1e-11)  # 4500 sequential rotations: rounding accumulates differently in fused form


def test_id_reuse_and_noncontiguous_ids(Backend):
    rng = np.random.default_rng(4)
    gpu, chk = Backend(2), checker(2)
    ids = [7, 1000000, 3, 4000000000]
    for q in ids:
        gpu.allocate_qubit(q)
        chk.allocate_qubit(q)
    for g in range(20):
        qs = [int(x) for x in rng.permutation(ids)[:3]]
        m = rand_unitary(rng, 2)
        gpu.apply_controlled_gate(m, qs[:2], qs[2:])
        chk.apply_controlled_gate(marg(chk, m), qs[:2], qs[2:])
    assert list(gpu.measure_qubits([3])) == list(chk.measure_qubits([3]))
    gpu.deallocate_qubit(3)
    chk.deallocate_qubit(3)
    gpu.allocate_qubit(3)
    chk.allocate_qubit(3)
    m = rand_unitary(rng, 2)
    gpu.apply_controlled_gate(m, [3, 7], [])
    chk.apply_controlled_gate(marg(chk, m), [3, 7], [])
    same(gpu, chk)


def test_sliced_pass_equals_whole_pass(Backend):
    """a dense pass restricted in turn to every slice of the state (the way the sharded engine runs the passes around a
    remap) equals one unrestricted pass"""
    rng = np.random.default_rng(77)
    n = 16
    for k, positions, ctrl_mask, slice_mask in [(1, [5], 0, 0b11 << 14), (2, [0, 9], 1 << 3, (1 << 15) | (1 << 1)),
                                                (4, [2, 3, 8, 11], 0, (1 << 15) | (1 << 14) | (1 << 13)),
                                                (4, [0, 1, 2, 3], 1 << 7, (1 << 4) | (1 << 12)),
                                                (3, [0, 1, 2], 0, (1 << 10) | (1 << 9) | (1 << 15)),
                                                (5, [1, 4, 6, 7, 13], 1 << 0, (1 << 2) | (1 << 15))]:
        gpu, chk = Backend(1), checker(1)
        for q in range(n):
            gpu.allocate_qubit(q)
            chk.allocate_qubit(q)
        wf = rand_state(rng, n)
        gpu.set_wavefunction(wf, list(range(n)))
        chk.set_wavefunction(warg(chk, wf), list(range(n)))
        m = rand_unitary(rng, k)
        gpu.selftest_sliced_pass(m, positions, ctrl_mask, slice_mask)
        chk.apply_controlled_gate(marg(chk, m), positions, [b for b in range(n) if (ctrl_mask >> b) & 1])
        chk.run()
        same(gpu, chk)


def test_exchange_kernel_permutation_single_device():
    """the peer-memory exchange kernel's index arithmetic (sub-block patterns, pair split, slices, 1-3 exchanged bits) on
    one device: `world` shards in one process, compared with the permutation a global<->local remap must perform"""
    from projectq_b200.backend import selftest_exchange

    for world, n_local, pairs, slice_mask in [(2, 10, [(0, 9)], 0), (2, 10, [(0, 0)], 0b110), (2, 12, [(0, 5)], (1 << 11) | (1 << 2)),
                                              (4, 12, [(0, 11), (1, 10)], 0), (4, 12, [(1, 3), (0, 7)], (1 << 9) | (1 << 8) | (1 << 0)),
                                              (8, 14, [(0, 13), (1, 12), (2, 11)], (1 << 10) | (1 << 9) | (1 << 8)),
                                              (8, 14, [(2, 4), (0, 6), (1, 13)], 1 << 12), (8, 13, [(1, 12)], 0),
                                              (4, 4, [(0, 3), (1, 2)], 0b01), (2, 1, [(0, 0)], 0)]:
        assert selftest_exchange(0, world, n_local, pairs, slice_mask) == 0, (world, n_local, pairs, slice_mask)


def test_checkpoint_roundtrip_and_zero_copy_view(Backend, tmp_path):
    """f3: save_state / load_state restore amplitudes (bit for bit), the qubit map (also after deallocations of non-top
    qubits, i.e. a non-identity physical layout) and the measurement RNG position; the CUDA-array-interface view aliases
    the state in HBM"""
    import torch

    from projectq_b200.backend import DeviceStateView

    rng = np.random.default_rng(8)
    n = 12
    gpu, chk = Backend(21), checker(21)
    ids = [3, 40, 7, 1, 9, 22, 5, 11, 2, 8, 30, 6]
    for q in ids:
        gpu.allocate_qubit(q)
        chk.allocate_qubit(q)
    wf = rand_state(rng, n)
    gpu.set_wavefunction(wf, ids)
    chk.set_wavefunction(warg(chk, wf), ids)
    for sim in (gpu, chk):  # one draw of the RNG stream, and a deallocation that leaves a hole in the layout
        bits = sim.measure_qubits([7])
        sim.deallocate_qubit(7)
    same(gpu, chk)
    prefix = str(tmp_path / "ckpt")
    gpu.save_state(prefix)
    m = rand_unitary(rng, 2)
    gpu.apply_controlled_gate(m, [3, 40], [])
    gpu.measure_qubits([1])  # moves the RNG on and changes the state
    fresh = Backend(999)
    fresh.load_state(prefix)
    same(fresh, chk)
    assert list(fresh.measure_qubits([9, 22])) == list(chk.measure_qubits([9, 22]))  # same RNG position, same state
    same(fresh, chk)
    with pytest.raises(RuntimeError):
        Backend(1).load_state(str(tmp_path / "missing"))
    # zero-copy view: writing through the tensor changes what the backend sees
    view = DeviceStateView(fresh)
    t = torch.as_tensor(view, device="cuda")
    assert t.dtype == torch.complex128 and t.numel() == 1 << fresh.num_qubits()
    before = np.asarray(fresh.cheat()[1]).copy()
    t.mul_(2.0)
    torch.cuda.synchronize()
    assert np.array_equal(np.asarray(fresh.cheat()[1]), 2.0 * before)
    assert sorted(view.layout) == list(range(fresh.num_qubits()))


def test_generic_math_function_is_called_on_every_register_value_like_the_reference(Backend):
    """The reference calls the Python function of a generic math gate for EVERY control-satisfying basis state, populated
    or not (simulator.hpp:246-253 has no amplitude test; _cppsim.cpp:33-41), so a function that raises on some register
    value raises whatever the state is.  The shim evaluates the function once per register value instead of once per basis
    state — the same set of arguments — so it raises for the same functions, and for total functions the results are
    identical."""
    ref = load_ref_cppsim()
    seen = set()

    def picky(x):
        seen.add(x[0])
        if x[0] == 5:
            raise ValueError("unreachable input")
        return [(x[0] + 1) % 8]

    sims = [Backend(1)] + ([ref.Simulator(1)] if ref is not None else [])
    for sim in sims:
        seen.clear()
        for q in range(4):
            sim.allocate_qubit(q)
        # |0000>: register value 5 is not populated, and still the function is asked about it
        with pytest.raises(Exception):
            sim.emulate_math(picky, [[0, 1, 2]], [])
        assert 5 in seen
    gpu, chk = Backend(1), checker(1)
    for q in range(5):
        gpu.allocate_qubit(q)
        chk.allocate_qubit(q)
    wf = rand_state(np.random.default_rng(2), 5)
    gpu.set_wavefunction(wf, list(range(5)))
    chk.set_wavefunction(warg(chk, wf), list(range(5)))
    f = lambda x: [(3 * x[0] + 1) % 8, x[1] ^ 1]  # noqa: E731
    gpu.emulate_math(f, [[0, 1, 2], [4]], [3])
    chk.emulate_math(f, [[0, 1, 2], [4]], [3])
    m1, v1 = gpu.cheat()
    assert np.array_equal(np.asarray(v1), np.asarray(chk.cheat()[1]))


_OPT_IN_SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, %r)
from oracle.statevec_oracle import OracleSimulator
from projectq_b200.backend import SimulatorBackend
from projectq_b200.workloads import tfim_terms
from tests.helpers import brickwork_circuit, rand_unitary

n = 16
gpu, chk = SimulatorBackend(3), OracleSimulator(3)
for q in range(n):
    gpu.allocate_qubit(q)
    chk.allocate_qubit(q)
rng = np.random.default_rng(5)
for m, t, c in brickwork_circuit(n, 3, seed=11):
    gpu.apply_controlled_gate(m, t, c)
    chk.apply_controlled_gate(m, t, c)
for pos in ([3, 7, 9, 12], [4, 5, 6, 15], [8, 10, 11, 13]):   # k = 4 with the three lowest bits free
    u = rand_unitary(rng, 4)
    gpu.apply_controlled_gate(u, pos, [14] if 14 not in pos else [])
    chk.apply_controlled_gate(u, pos, [14] if 14 not in pos else [])
gpu.run()
chk.run()
terms = tfim_terms(n) + [([(0, "Y"), (9, "Z"), (15, "X")], 0.3), ([(12, "Z"), (13, "Z"), (2, "Z")], -0.2)]
ids = list(range(n))
e_gpu, e_chk = gpu.get_expectation_value(terms, ids), chk.get_expectation_value(terms, ids)
assert abs(e_gpu - e_chk) < 1e-11, (e_gpu, e_chk)
gpu.emulate_time_evolution(terms, 0.07, ids, [])
chk.emulate_time_evolution(terms, 0.07, ids, [])
cterms = [(t, c * (0.6 + 0.2j)) for t, c in terms]
gpu.apply_qubit_operator(cterms, ids)
chk.apply_qubit_operator(cterms, ids)
err = float(np.max(np.abs(np.asarray(gpu.cheat()[1]) - chk.cheat()[1])))
assert err < 1e-11, err
print("opt-in paths OK", err)
"""


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"PQB_PAULI_BLOCK_BITS": "13"}, {"PQB_PAULI_BLOCK_BITS": "30"}, {"PQB_DENSE_DMMA": "1"},
                                 {"PQB_DENSE_DMMA": "2"}, {"PQB_PAULI_CTAS_PER_SM": "3"}])
def test_opt_in_kernel_paths_match_the_oracle(env):
    """the measured-and-not-adopted kernel forms stay correct: tile-bit sets fused block by block (several blocks / one block),
    the k = 4 pass on the FP64 tensor pipe (register and cp.async-ring form), the 3-CTA Pauli tile kernel"""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", _OPT_IN_SCRIPT % root], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, **env), cwd=root)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "opt-in paths OK" in out.stdout


def _numpy_math(vec, n, regs, ctrl, f):
    """new[pi(i)] += psi[i] (simulator.hpp:224-269) vectorised, 64-bit register arithmetic, sources added in index order"""
    idx = np.arange(1 << n, dtype=np.int64)
    cmask = sum(1 << c for c in ctrl)
    ni = idx.copy()
    for reg in regs:
        x = np.zeros_like(idx)
        for b, p in enumerate(reg):
            x |= ((idx >> p) & 1) << b
        y = f(x) & ((1 << len(reg)) - 1)
        for b, p in enumerate(reg):
            ni = (ni & ~(1 << p)) | (((y >> b) & 1) << p)
    ni = np.where((idx & cmask) == cmask, ni, idx)
    new = np.zeros_like(vec)
    np.add.at(new, ni, vec)
    return new


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["add", "addmod_near", "addmod_small_N", "mulmod", "mulmod_two_regs", "mulmod_not_coprime"])
def test_emulate_math_wide_registers_closed_form_inverse(Backend, case):
    """registers too wide to tabulate (> 20 bits): the closed-form inverse gather, bit-exact against a NumPy scatter — also
    where the random state populates values outside the gate's domain (x >= N: several sources per destination), and the
    atomic fallback when a is not invertible mod N"""
    n = 24
    rng = np.random.default_rng(17)
    wf = rand_state(rng, n)
    gpu = Backend(1)
    for q in range(n):
        gpu.allocate_qubit(q)
    gpu.set_wavefunction(wf, list(range(n)))
    scrambled = [3, 1, 2, 4, 5, 6, 7, 9, 8, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22]  # 22 bits, three runs
    if case == "add":
        regs, ctrl, args = [scrambled], [23], ("emulate_math_addConstant", (-1234567,))
        f = lambda x: x - 1234567
    elif case == "addmod_near":
        N = 4194301
        regs, ctrl, args = [scrambled], [0], ("emulate_math_addConstantModN", (4000000, N))
        f = lambda x: (x + 4000000) % N
    elif case == "addmod_small_N":
        N = 1500007
        regs, ctrl, args = [list(range(1, 23))], [], ("emulate_math_addConstantModN", (77, N))
        f = lambda x: (x + 77) % N
    elif case == "mulmod":
        N = 4194301
        regs, ctrl, args = [list(range(2, 24))], [0, 1], ("emulate_math_multiplyByConstantModN", (1234577, N))
        f = lambda x: (x * 1234577) % N
    elif case == "mulmod_two_regs":
        N = 2039
        regs, ctrl, args = [list(range(0, 11)), list(range(12, 23))], [23], ("emulate_math_multiplyByConstantModN", (7, N))
        f = lambda x: (x * 7) % N
    else:
        N = 4194300  # gcd(6, N) != 1: not a permutation, the scatter with atomics takes it
        regs, ctrl, args = [list(range(1, 23))], [], ("emulate_math_multiplyByConstantModN", (6, N))
        f = lambda x: (x * 6) % N
    getattr(gpu, args[0])(*args[1], regs, ctrl)
    got = np.asarray(gpu.cheat()[1])
    want = _numpy_math(wf, n, regs, ctrl, f)
    if case == "mulmod_not_coprime":
        assert np.max(np.abs(got - want)) < 1e-12  # atomics: the order of the additions is not fixed
    else:
        assert np.array_equal(got, want), float(np.max(np.abs(got - want)))
