"""Wall time of two single-qubit measurements at 30 qubits (for an ncu launch list).  Not part of the product path."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projectq_b200.backend import SimulatorBackend
sim = SimulatorBackend(1)
sim.init_random_state(30, 42)
for q in (7, 19):
    sim.synchronize(); t0 = time.perf_counter()
    b = sim.measure_qubits([q])
    sim.synchronize(); print("measure", q, b, (time.perf_counter() - t0) * 1e3, "ms", flush=True)
