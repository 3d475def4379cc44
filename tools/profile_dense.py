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
for pos in ([8, 9, 10, 11], [3, 4, 5, 6], [20, 23, 26, 29], [5, 12, 19, 27]):
    ms = sim.bench_dense_pass(rand_unitary(rng, 4), pos, 0, 3)
    print("dense k=4 %s at %dq: %.3f ms, %.0f GB/s" % (pos, n, ms, 32.0 * (1 << n) / ms / 1e6), flush=True)
