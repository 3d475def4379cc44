"""Synthetic workloads of the BASELINE configs and their on-disk formats (shared by bench.py, tools/ and the tests).

* ``brickwork_circuit`` — BASELINE config 2 / 5 generator.
* ``tfim_terms`` — the open-chain transverse-field Ising Hamiltonian of config 4 in the native seam's term format.
* ``pack_gate_stream`` — the packed gate list that ``pqb_apply_gate_stream`` ingests in one call.
* ``write_circuit_file`` / ``write_ops_file`` — input files of the reference harness (oracle/ref_harness.cpp), used by the
  parity tests and the CPU baseline to drive the unmodified reference C++ simulator on exactly the same workload.
"""
import struct

import numpy as np


def brickwork_circuit(n, depth, seed=2026):
    """Per layer one of Rx/Ry/Rz(theta) on every qubit, then CNOT or CZ on (q, q+1) for q in range(d % 2, n - 1, 2).
    Returns [(matrix, targets, ctrls)] as the Simulator receives them (CNOT = X with a control, CZ = Z with a control)."""
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


def inverse_circuit(gates):
    """U^dagger of a gate list: reversed order, conjugate-transposed matrices, same targets and controls."""
    return [(np.ascontiguousarray(m.conj().T), t, c) for m, t, c in reversed(gates)]


def ry_layer(n, seed=7):
    """product-state preparation of config 4: one Ry(theta_q) per qubit"""
    rng = np.random.default_rng(seed)
    gates = []
    for q in range(n):
        th = float(rng.uniform(0, np.pi))
        c, s = np.cos(th / 2), np.sin(th / 2)
        gates.append((np.array([[c, -s], [s, c]], dtype=np.complex128), [q], []))
    return gates


def tfim_terms(n, J=1.0, h=0.7):
    terms = [([(i, "Z"), (i + 1, "Z")], -J) for i in range(n - 1)]
    terms += [([(i, "X")], -h) for i in range(n)]
    return terms


def pack_gate_stream(gates):
    """gates: list of (matrix ndarray 2^k x 2^k, targets, ctrls) -> (bytes, n); per gate u32 k, u32 nc, u32 targets[k],
    u32 ctrls[nc], f64 matrix[2*4^k] (the layout of include/pqb200.h pqb_apply_gate_stream and of the harness file)."""
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


def _ids(v):
    return struct.pack("<I", len(v)) + np.asarray(list(v), dtype=np.uint32).tobytes()


def _terms(terms):
    out = struct.pack("<I", len(terms))
    for term, coeff in terms:
        out += struct.pack("<dI", float(coeff), len(term))
        for idx, p in term:
            out += struct.pack("<II", int(idx), ord(p))
    return out


def write_ops_file(path, ops):
    """ops: list of tuples understood by oracle/ref_harness.cpp —
    ("evolve", time, ids, ctrl, terms) | ("expectation", ids, terms) | ("probability", ids, bits) | ("measure", ids) |
    ("mulmod", a, N, ids, ctrl).  Returns the number of f64 results the harness will write."""
    body = b""
    n_results = 0
    for op in ops:
        kind = op[0]
        if kind == "evolve":
            _, t, ids, ctrl, terms = op
            body += struct.pack("<Id", 1, float(t)) + _ids(ids) + _ids(ctrl) + _terms(terms)
        elif kind == "expectation":
            _, ids, terms = op
            body += struct.pack("<I", 2) + _ids(ids) + _terms(terms)
            n_results += 1
        elif kind == "probability":
            _, ids, bits = op
            body += struct.pack("<I", 3) + _ids(ids) + np.asarray([int(b) for b in bits], dtype=np.uint32).tobytes()
            n_results += 1
        elif kind == "measure":
            _, ids = op
            body += struct.pack("<I", 4) + _ids(ids)
            n_results += len(ids)
        elif kind == "mulmod":
            _, a, N, ids, ctrl = op
            body += struct.pack("<Iqq", 5, int(a), int(N)) + _ids(ids) + _ids(ctrl)
        else:
            raise ValueError("unknown op %r" % (kind,))
    with open(path, "wb") as f:
        f.write(b"PQBO" + struct.pack("<II", 1, len(ops)) + body)
    return n_results
