"""Single-GPU fixtures for the sharded runs' parity block (run on a 1-GPU B200 box):

    python tools/make_sharded_fixture.py [n ...]        -> gpurun_out/brickwork_1gpu_<n>q.json (copy to tests/golden/)

For every n: the seeded random state of bench.py (init_random_state(n, 2026)), ONE step of the config-2 circuit on ONE GPU,
amplitudes at 256 seeded indices.  bench.py at N = 2, 4, 8 (n = 31, 32, 33; and the 33-qubit strong-scaling leg at every N)
must reproduce these amplitudes within 1e-12 on the sharded state.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import DEPTH, N_PARITY_SAMPLES, STATE_SEED  # noqa: E402
from projectq_b200.backend import SimulatorBackend  # noqa: E402
from projectq_b200.workloads import brickwork_circuit, pack_gate_stream  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [30, 31, 32, 33]
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    for n in sizes:
        gates = brickwork_circuit(n, DEPTH)
        body, n_gates = pack_gate_stream(gates)
        idx = np.random.default_rng(n).integers(0, 1 << n, N_PARITY_SAMPLES, dtype=np.uint64)
        sim = SimulatorBackend(1)
        sim.init_random_state(n, STATE_SEED)
        sim.apply_gate_stream(body, n_gates, True)
        sim.run()
        amps = np.asarray(sim.get_amplitudes(idx))
        norm = sim.norm_squared()
        del sim
        path = os.path.join(out_dir, "brickwork_1gpu_%dq.json" % n)
        with open(path, "w") as f:
            json.dump({"config": "config-2 brickwork circuit, %d qubits, depth %d, one step from init_random_state(n, %d) on one "
                                 "B200 (tools/make_sharded_fixture.py)" % (n, DEPTH, STATE_SEED),
                       "indices": [int(i) for i in idx], "amplitudes": [[a.real, a.imag] for a in amps], "norm": norm}, f)
        print("wrote", path, "norm", norm, flush=True)


if __name__ == "__main__":
    main()
