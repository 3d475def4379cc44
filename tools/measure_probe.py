import sys, time
sys.path.insert(0, "/root/repo")
from projectq_b200.backend import SimulatorBackend
sim = SimulatorBackend(1)
sim.init_random_state(30, 42)
for q in (7, 19):
    sim.synchronize(); t0 = time.perf_counter()
    b = sim.measure_qubits([q])
    sim.synchronize(); print("measure", q, b, (time.perf_counter() - t0) * 1e3, "ms", flush=True)
