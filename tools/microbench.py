"""Per-kernel micro-benchmarks on the GPU box: dense k-qubit pass time by target placement, reductions, peaks.
Writes JSON lines to stdout (and optionally to a file).  Not part of the product path."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projectq_b200.backend import SimulatorBackend  # noqa: E402
from tests.helpers import rand_unitary  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else None
    sim = SimulatorBackend(1)

    def emit(d):
        line = json.dumps(d)
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()

    emit({"fp64_tflops": sim.measure_fp64_peak(), "copy_gbs": sim.measure_copy_bandwidth(1 << 30)})
    sim.init_random_state(n, 42)
    rng = np.random.default_rng(0)
    amps = 1 << n
    for k in range(1, 6):
        m = rand_unitary(rng, k)
        placements = {
            "low": list(range(k)),
            "mid": list(range(8, 8 + k)),
            "high": list(range(n - k, n)),
            "spread": sorted(int(x) for x in np.linspace(0, n - 1, k).round()),
            "low1_high": [0] + list(range(n - k + 1, n)),
            "low2_high": [0, 1] + list(range(n - k + 2, n)) if k >= 3 else [0, 0],
            "low3_high": [0, 1, 2] + list(range(n - k + 3, n)) if k >= 4 else [0, 0],
        }
        for name, pos in placements.items():
            if len(pos) != k or len(set(pos)) != k:
                continue
            ms = sim.bench_dense_pass(m, pos, 0, 5)
            emit({"k": k, "placement": name, "pos": pos, "ms": ms, "GBs": 32.0 * amps / ms / 1e6,
                  "fp64_tflops": (1 << k) * 8.0 * amps / ms / 1e9})
        ms = sim.bench_dense_pass(m, list(range(4, 4 + k)), 1 << (n - 1), 5)
        emit({"k": k, "placement": "mid+ctrl_top", "ms": ms, "GBs": 16.0 * amps / ms / 1e6})
    sim.timer_start()
    for _ in range(5):
        sim.norm_squared()
    ms = sim.timer_stop() / 5
    emit({"op": "norm_squared", "ms": ms, "GBs": 16.0 * amps / ms / 1e6})


if __name__ == "__main__":
    main()
