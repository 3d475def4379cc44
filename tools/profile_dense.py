"""One k = 4 pass at 30 qubits per placement, for ncu (kernel selection by environment, e.g. PQB_DENSE_DMMA).  Not product path."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projectq_b200.backend import SimulatorBackend  # noqa: E402
from tests.helpers import rand_unitary  # noqa: E402

rng = np.random.default_rng(0)
n = int(os.environ.get("PQB_PROFILE_QUBITS", "30"))
sim = SimulatorBackend(1)
sim.init_random_state(n, 42)
k = int(os.environ.get("PQB_PROFILE_K", "4"))
placements = {4: ([8, 9, 10, 11], [3, 4, 5, 6], [20, 23, 26, 29], [5, 12, 19, 27]),
              5: ([8, 9, 10, 11, 12], [18, 20, 23, 26, 29], [2, 9, 10, 11, 12], [0, 1, 2, 3, 4], [0, 5, 6, 7, 8], [1, 2, 3, 4, 5], [0, 2, 4, 6, 8])}[k]
for pos in placements:
    ms = sim.bench_dense_pass(rand_unitary(rng, k), pos, 0, 3)
    print("dense k=%d %s at %dq: %.3f ms, %.0f GB/s, %.1f TFLOP/s" %
          (k, pos, n, ms, 32.0 * (1 << n) / ms / 1e6, 8.0 * (1 << k) * (1 << n) / ms / 1e9), flush=True)
