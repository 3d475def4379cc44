"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np


def rand_unitary(rng, k):
    d = 1 << k
    a = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def rand_state(rng, n):
    wf = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return wf / np.linalg.norm(wf)


def fidelity_gap(a, b):
    return abs(1.0 - abs(np.vdot(a, b)) ** 2 / (np.vdot(a, a).real * np.vdot(b, b).real))


def tfim_terms(n, J=1.0, h=0.7):
    terms = [([(i, "Z"), (i + 1, "Z")], -J) for i in range(n - 1)]
    terms += [([(i, "X")], -h) for i in range(n)]
    return terms


def pack_gate_stream(gates):
    """gates: list of (matrix ndarray 2^k x 2^k, targets, ctrls) -> (bytes, n) in the layout of oracle/ref_harness.cpp."""
    out = bytearray()
    for m, t, c in gates:
        out += np.array([len(t), len(c)], dtype=np.uint32).tobytes()
        out += np.array(list(t) + list(c), dtype=np.uint32).tobytes()
        out += np.ascontiguousarray(m, dtype=np.complex128).tobytes()
    return bytes(out), len(gates)


def write_circuit_file(path, n_qubits, gates):
    body, n = pack_gate_stream(gates)
    with open(path, "wb") as f:
        f.write(b"PQBC")
        f.write(np.array([1, n_qubits, n], dtype=np.uint32).tobytes())
        f.write(body)


def brickwork_circuit(n, depth, seed=2026):
    """BASELINE config 2 generator: per layer one of Rx/Ry/Rz(theta) on every qubit, then CNOT or CZ on
    (q, q+1) for q in range(d % 2, n - 1, 2).  Returns [(matrix, targets, ctrls)] as the Simulator would receive them
    (CNOT = X with a control, CZ = Z with a control)."""
    rng = np.random.default_rng(seed)
    gates = []
    X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
    for d in range(depth):
        for q in range(n):
            kind = int(rng.integers(0, 3))
            th = float(rng.uniform(0, 2 * np.pi))
            c, s = np.cos(th / 2), np.sin(th / 2)
            if kind == 0:
                m = np.array([[c, -1j * s], [-1j * s, c]])
            elif kind == 1:
                m = np.array([[c, -s], [s, c]], dtype=np.complex128)
            else:
                m = np.array([[np.exp(-0.5j * th), 0], [0, np.exp(0.5j * th)]])
            gates.append((m.astype(np.complex128), [q], []))
        for q in range(d % 2, n - 1, 2):
            gates.append((X if int(rng.integers(0, 2)) == 0 else Z, [q + 1], [q]))
    return gates
