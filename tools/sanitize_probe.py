"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck).  Not part of the product path.

    compute-sanitizer --tool memcheck  python tools/sanitize_probe.py
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.statevec_oracle import OracleSimulator  # noqa: E402  (checker only)
from projectq_b200.backend import SimulatorBackend  # noqa: E402
from projectq_b200.workloads import brickwork_circuit, tfim_terms  # noqa: E402
from tests.helpers import rand_unitary  # noqa: E402

n = int(os.environ.get("PQB_PROBE_QUBITS", "15"))
rng = np.random.default_rng(3)
gpu, chk = SimulatorBackend(5), OracleSimulator(5)
for q in range(n):
    gpu.allocate_qubit(q)
    chk.allocate_qubit(q)
for m, t, c in brickwork_circuit(n, 2, seed=4):
    gpu.apply_controlled_gate(m, t, c)
    chk.apply_controlled_gate(m, t, c)
for k, pos, ctrl in ((5, [5, 6, 8, 9, 12], [13]), (5, [0, 1, 2, 3, 4], []), (5, [1, 3, 7, 10, 14], [0]), (4, [0, 1, 2, 3], [9]),
                     (4, [4, 6, 9, 11], []), (3, [0, 1, 2], []), (2, [0, 13], [5, 6]), (1, [7], [])):
    u = rand_unitary(rng, k)
    gpu.apply_controlled_gate(u, pos, ctrl)
    gpu.run()
    chk.apply_controlled_gate(u, pos, ctrl)
    chk.run()
terms = tfim_terms(n) + [([(0, "Y"), (9, "Z"), (14, "X")], 0.3), ([(12, "Z"), (13, "Z"), (2, "Z")], -0.2)]
ids = list(range(n))
e1, e2 = gpu.get_expectation_value(terms, ids), chk.get_expectation_value(terms, ids)
assert abs(e1 - e2) < 1e-11, (e1, e2)
gpu.emulate_time_evolution(terms, 0.05, ids, [3])
chk.emulate_time_evolution(terms, 0.05, ids, [3])
gpu.apply_qubit_operator([(t, c * (0.5 - 0.1j)) for t, c in terms], ids)
chk.apply_qubit_operator([(t, c * (0.5 - 0.1j)) for t, c in terms], ids)
err = float(np.max(np.abs(np.asarray(gpu.cheat()[1]) - chk.cheat()[1])))
assert err < 1e-11, err
nrm = float(np.linalg.norm(chk.cheat()[1]))
wf = chk.cheat()[1] / nrm
gpu.set_wavefunction(wf, ids)
chk.set_wavefunction(wf, ids)
gpu.emulate_math_multiplyByConstantModN(5, 123, [list(range(1, 8))], [9])
chk.emulate_math_multiplyByConstantModN(5, 123, [list(range(1, 8))], [9])
gpu.emulate_math_addConstant(7, [list(range(0, 6)), list(range(8, 13))], [])
chk.emulate_math_addConstant(7, [list(range(0, 6)), list(range(8, 13))], [])
assert np.array_equal(np.asarray(gpu.cheat()[1]), chk.cheat()[1])
p1, p2 = gpu.get_probability([1, 0], [2, 11]), chk.get_probability([1, 0], [2, 11])
assert abs(p1 - p2) < 1e-12
b1, b2 = list(gpu.measure_qubits([3, 8, 14])), list(chk.measure_qubits([3, 8, 14]))
assert b1 == b2, (b1, b2)
gpu.deallocate_qubit(14) if gpu.is_classical(14, 1e-10) else None
print("sanitize probe OK", err, b1)
