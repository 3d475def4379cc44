"""One launch of each hot kernel at BASELINE sizes, for `ncu --set full` (one GPU).  Not part of the product path.

    ncu --set full --clock-control none --import-source on \
        -k regex:"apply_dense_kernel|pauli_tile_kernel|emulate_math_gather_kernel|norm_masked_kernel" \
        -o gpurun_out/r2_prof_kernels python tools/profile_kernels.py

Prints the CUDA-event / wall time of every operation as well (taken outside the profiler when run without ncu).
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projectq_b200.backend import SimulatorBackend  # noqa: E402
from projectq_b200.workloads import tfim_terms  # noqa: E402
from tests.helpers import rand_unitary  # noqa: E402

rng = np.random.default_rng(0)


def wall(sim, fn):
    sim.synchronize()
    t0 = time.perf_counter()
    fn()
    sim.synchronize()
    return (time.perf_counter() - t0) * 1e3


# ---- config 2: the dominant kernel, one k = 4 pass at 30 qubits (mid placement) + one on the lowest bits ----
n = 30
sim = SimulatorBackend(1)
sim.init_random_state(n, 42)
for pos in ([8, 9, 10, 11], [0, 1, 2, 3]):
    ms = sim.bench_dense_pass(rand_unitary(rng, 4), pos, 0, 1)
    print("dense k=4 %s at 30q: %.3f ms, %.0f GB/s" % (pos, ms, 32.0 * (1 << n) / ms / 1e6), flush=True)
# ---- config 3: the permutation kernel, controlled (x * 7) mod 4087 on a 12-bit register inside 30 qubits ----
sim.emulate_math_multiplyByConstantModN(7, 4087, [list(range(4, 16))], [29])  # allocates the scratch copy
ms = wall(sim, lambda: sim.emulate_math_multiplyByConstantModN(7, 4087, [list(range(4, 16))], [29]))
print("emulate_math mul mod N at 30q: %.3f ms, %.0f GB/s of 32 B/amp" % (ms, 32.0 * (1 << n) / ms / 1e6), flush=True)
ms = wall(sim, lambda: sim.norm_squared())
print("norm_squared at 30q: %.3f ms, %.0f GB/s of 16 B/amp" % (ms, 16.0 * (1 << n) / ms / 1e6), flush=True)
del sim

# ---- config 4: TFIM at 28 qubits: <H> and one application of H (= one Taylor order without the accumulation) ----
n = 28
sim = SimulatorBackend(1)
sim.init_random_state(n, 7)
terms = tfim_terms(n)
ids = list(range(n))
sim.get_expectation_value(terms, ids)
ms = wall(sim, lambda: sim.get_expectation_value(terms, ids))
print("TFIM-28 <H> (55 terms): %.3f ms = %.2f sweeps of the state at the measured HBM peak" % (ms, ms / (16.0 * (1 << n) / 6546.6e6)),
      flush=True)
cterms = [(t, complex(c)) for t, c in terms]
sim.apply_qubit_operator(cterms, ids)
ms = wall(sim, lambda: sim.apply_qubit_operator(cterms, ids))
print("TFIM-28 apply_qubit_operator (55 terms): %.3f ms, %.1f B/amp-equivalents at the measured HBM peak" %
      (ms, ms * 6546.6e6 / (1 << n)), flush=True)
