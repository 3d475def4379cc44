"""Where the time of the config-1 / config-3 call-trace replays goes, per native-seam method (wall time around each call)."""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from projectq_b200.backend import SimulatorBackend  # noqa: E402
from tests.test_golden import dec, load  # noqa: E402


def replay(sim, trace, acc=None):
    t0 = time.perf_counter()
    for e in trace:
        if e["m"].startswith("_"):
            continue
        args = dec(e["a"])
        if e["m"] == "apply_controlled_gate":
            args[0] = np.array(args[0], dtype=np.complex128)
        t1 = time.perf_counter()
        getattr(sim, e["m"])(*args)
        if acc is not None:
            acc[e["m"]][0] += time.perf_counter() - t1
            acc[e["m"]][1] += 1
    sim.synchronize()
    return time.perf_counter() - t0


for name, seed in (("qft20", 1), ("shor4087", 3)):
    data = load(name)
    replay(SimulatorBackend(seed), data["trace"])
    ts = [replay(SimulatorBackend(seed), data["trace"]) for _ in range(8)]
    acc = collections.defaultdict(lambda: [0.0, 0])
    total = replay(SimulatorBackend(seed), data["trace"], acc)
    print(name, "runs (ms):", [round(t * 1e3, 2) for t in ts], "; one run by method (ms, calls):",
          {k: (round(v[0] * 1e3, 2), v[1]) for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0])}, "total %.2f" % (total * 1e3), flush=True)
