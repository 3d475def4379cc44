"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from projectq_b200.workloads import (brickwork_circuit, inverse_circuit, pack_gate_stream, ry_layer, tfim_terms,  # noqa: F401
                                    write_circuit_file, write_ops_file)


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
