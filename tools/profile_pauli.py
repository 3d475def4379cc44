"""Config 4's kernels alone at 28 qubits (TFIM, 55 terms): <H>, one application of H, one time evolution.  For ncu:

    ncu --set full --clock-control none --import-source on -k regex:pauli_tile -c 6 \
        -o gpurun_out/r2_prof_pauli python tools/profile_pauli.py --no-evolution

Without a profiler it prints wall times (the engine's stream is synchronised on both sides).  Not part of the product path.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projectq_b200.backend import SimulatorBackend  # noqa: E402
from projectq_b200.workloads import tfim_terms  # noqa: E402


def wall(sim, fn, reps=1):
    sim.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    sim.synchronize()
    return (time.perf_counter() - t0) * 1e3 / reps


n = int(os.environ.get("PQB_PROFILE_QUBITS", "28"))
sim = SimulatorBackend(1)
sim.init_random_state(n, 7)
terms = tfim_terms(n)
ids = list(range(n))
reps = 1 if "--no-evolution" in sys.argv else 5
sim.get_expectation_value(terms, ids)
ms = wall(sim, lambda: sim.get_expectation_value(terms, ids), reps)
print("TFIM-%d <H> (%d terms): %.3f ms = %.2f sweeps of the state at the measured HBM peak" %
      (n, len(terms), ms, ms / (16.0 * (1 << n) / 6546.6e6)), flush=True)
cterms = [(t, complex(c)) for t, c in terms]
sim.apply_qubit_operator(cterms, ids)
ms = wall(sim, lambda: sim.apply_qubit_operator(cterms, ids), reps)
print("TFIM-%d apply_qubit_operator (%d terms): %.3f ms, %.1f B/amp-equivalents at the measured HBM peak" %
      (n, len(terms), ms, ms * 6546.6e6 / (1 << n)), flush=True)
if "--no-evolution" not in sys.argv:
    del sim
    sim = SimulatorBackend(1)
    sim.init_random_state(n, 7)
    sim.emulate_time_evolution(terms, 0.1, ids, [])
    l0 = sim.stats()["kernel_launches"]
    ms = wall(sim, lambda: sim.emulate_time_evolution(terms, 0.1, ids, []))
    print("TFIM-%d emulate_time_evolution(t=0.1): %.1f ms, %d kernel launches" % (n, ms, sim.stats()["kernel_launches"] - l0), flush=True)
