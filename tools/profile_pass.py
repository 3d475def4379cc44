"""A handful of dense passes on a 28-qubit state for `ncu --set full` (one GPU).  Not part of the product path."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projectq_b200.backend import SimulatorBackend  # noqa: E402
from tests.helpers import rand_unitary  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
sim = SimulatorBackend(1)
sim.init_random_state(n, 42)
rng = np.random.default_rng(0)
# (k, positions): mid / low placements of the widths the fuser emits
for k, pos in ((4, [8, 9, 10, 11]), (4, [0, 1, 2, 3]), (5, [8, 9, 10, 11, 12]), (5, [0, 1, 2, 3, 4]), (3, [0, 1, 2]),
               (4, [5, 13, 20, 27])):
    ms = sim.bench_dense_pass(rand_unitary(rng, k), pos, 0, 1)
    print(k, pos, "%.3f ms" % ms, "%.0f GB/s" % (32.0 * (1 << n) / ms / 1e6))
