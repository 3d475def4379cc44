import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from projectq_b200.backend import SimulatorBackend
s = SimulatorBackend(1); del s
for rep in range(3):
    t0 = time.perf_counter(); s = SimulatorBackend(1); t1 = time.perf_counter()
    s.allocate_qubit(0); t2 = time.perf_counter()
    for q in range(1, 13): s.allocate_qubit(q)
    t3 = time.perf_counter()
    del s; t4 = time.perf_counter()
    print("ctor %.3f ms, first allocate %.3f ms, 12 more %.3f ms, dtor %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3), flush=True)
