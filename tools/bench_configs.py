"""Timings of BASELINE.json configs 1, 3 and 4 and of the non-gate kernels, with the reference C++ simulator (oracle/_ref)
timed beside them on the host.  Writes JSON lines.  Not part of the product path (test/measurement infrastructure)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from projectq_b200.backend import SimulatorBackend  # noqa: E402
from tests.conftest import load_ref_cppsim  # noqa: E402
from tests.helpers import tfim_terms  # noqa: E402
from tests.test_golden import dec, load  # noqa: E402

out = open(sys.argv[1], "w") if len(sys.argv) > 1 else None
ref = load_ref_cppsim()


def emit(d):
    line = json.dumps(d)
    print(line, flush=True)
    if out:
        out.write(line + "\n")
        out.flush()


def replay(sim, trace, as_list):
    t0 = time.perf_counter()
    for e in trace:
        if e["m"].startswith("_"):
            continue
        args = dec(e["a"])
        if e["m"] == "apply_controlled_gate" and not as_list:
            args[0] = np.array(args[0], dtype=np.complex128)
        getattr(sim, e["m"])(*args)
    if hasattr(sim, "synchronize"):
        sim.synchronize()
    return time.perf_counter() - t0


# ---- configs 1 and 3: replay of the reference engine's own call sequence at the native seam ----
for name, seed in (("qft20", 1), ("shor4087", 3)):
    data = load(name)
    replay(SimulatorBackend(seed), data["trace"], False)  # warm-up (module load, first launches)
    t_gpu = min(replay(SimulatorBackend(seed), data["trace"], False) for _ in range(3))
    t_ref = min(replay(ref.Simulator(seed), data["trace"], True) for _ in range(3)) if ref else None
    emit({"config": data["config"], "calls": len(data["trace"]), "gpu_s": t_gpu, "reference_cpu_s": t_ref,
          "cores": os.cpu_count()})

# ---- configs 1 and 3 literally as written: wall time through MainEngine (compiler chain + engine + native backend) ----
def through_main_engine():
    from tests import refenv

    if ref is not None:  # the compiled reference is already loaded (a pybind11 module cannot be loaded twice)
        sys.modules.setdefault("projectq.backends._sim._cppsim", ref)
    if refenv.import_projectq("reference") is None:
        emit({"config": "MainEngine legs", "unavailable": "reference front end not staged (baseline/_ref)"})
        return
    from projectq.backends import Simulator as RefSimulator

    from projectq_b200 import Simulator as CudaSimulator
    from tests.golden import programs

    def run(name, make_sim):
        t0 = time.perf_counter()
        result = None
        for step in programs.PROGRAMS[name](make_sim):
            if step[0] == "done":
                result = step[1]
        return time.perf_counter() - t0, result

    for name in ("qft20", "shor4087"):
        legs = {}
        for label, cls in (("reference", RefSimulator), ("cuda", CudaSimulator)):
            make = lambda gate_fusion, rnd_seed, cls=cls: cls(gate_fusion=gate_fusion, rnd_seed=rnd_seed)  # noqa: E731
            run(name, make)  # warm-up: decomposition rule caches, first launches
            legs[label] = min(run(name, make) for _ in range(3))
        (t_ref, r_ref), (t_gpu, r_gpu) = legs["reference"], legs["cuda"]
        emit({"config": name + " through MainEngine (program of tests/golden/programs.py), wall time incl. the Python compiler chain",
              "reference_engine_plus_cppsim_s": t_ref, "our_engine_plus_cuda_s": t_gpu, "speedup": t_ref / t_gpu,
              "same_measured_bits": r_ref == r_gpu})


through_main_engine()

# ---- config 4: TFIM VQE-style iteration ----
def tfim_iteration(sim, n, t_evolve, sync):
    terms = tfim_terms(n)
    ids = list(range(n))
    t0 = time.perf_counter()
    sim.emulate_time_evolution(terms, t_evolve, ids, [])
    if sync:
        sim.synchronize()
    t1 = time.perf_counter()
    e = sim.get_expectation_value(terms, ids)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, e


def ry_layer(sim, n, as_list):
    rng = np.random.default_rng(4)
    for q in range(n):
        th = float(rng.uniform(0, np.pi))
        m = np.array([[np.cos(th / 2), -np.sin(th / 2)], [np.sin(th / 2), np.cos(th / 2)]], dtype=np.complex128)
        sim.apply_controlled_gate(m.tolist() if as_list else m, [q], [])
    sim.run()


for n in (22, 28):
    sim = SimulatorBackend(1)
    for q in range(n):
        sim.allocate_qubit(q)
    ry_layer(sim, n, False)
    e0 = sim.get_expectation_value(tfim_terms(n), list(range(n)))
    tev, texp, e1 = tfim_iteration(sim, n, 0.1, True)
    sim.reset_stats()
    tev2, texp2, e2 = tfim_iteration(sim, n, 0.1, True)
    launches = sim.stats()["kernel_launches"]
    amps = float(1 << n)
    # t = 0.1, ||H||_1 = (n-1) + 0.7 n  ->  s = floor(0.1 * ||H||_1 + 1) sub-steps; the number of Taylor orders per sub-step is
    # data dependent (until the increment's norm drops below 1e-12): derived from the launch count of one evolution
    emit({"config": "TFIM %d qubits (%d terms): emulate_time_evolution(t=0.1) + get_expectation_value" % (n, 2 * n - 1),
          "gpu_time_evolution_s": min(tev, tev2), "gpu_expectation_s": min(texp, texp2), "E0": e0, "E1": e1, "E2": e2,
          "energy_drift": abs(e2 - e0), "norm": sim.norm_squared(), "kernel_launches_evolution_plus_expectation": launches,
          "floor_bytes_per_amp_per_taylor_order": 64, "floor_bytes_per_amp_expectation": 16,
          "expectation_sweeps_equivalent": min(texp, texp2) / (16.0 * amps / 6546.6e9)})
    del sim
    if n == 22 and ref:
        r = ref.Simulator(1)
        for q in range(n):
            r.allocate_qubit(q)
        ry_layer(r, n, True)
        t0 = time.perf_counter()
        e_ref0 = r.get_expectation_value(tfim_terms(n), list(range(n)))
        t1 = time.perf_counter()
        r.emulate_time_evolution(tfim_terms(n), 0.1, list(range(n)), [])
        t2 = time.perf_counter()
        e_ref1 = r.get_expectation_value(tfim_terms(n), list(range(n)))
        emit({"config": "reference C++ (host, %d threads), TFIM 22 qubits" % (os.cpu_count() or 1),
              "reference_time_evolution_s": t2 - t1, "reference_expectation_s": t1 - t0, "E0": e_ref0, "E1": e_ref1,
              "dE0_vs_gpu": abs(e_ref0 - e0), "dE1_vs_gpu": abs(e_ref1 - e1)})

# ---- non-gate kernels at 30 qubits ----
n = 30
sim = SimulatorBackend(5)
sim.init_random_state(n, 11)
amps = float(1 << n)


def timed(fn, reps=3):
    fn()
    sim.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        sim.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


t = timed(lambda: sim.norm_squared())
emit({"op": "norm_squared 30q (every amplitude read once)", "s": t, "GBs": 16 * amps / t / 1e9})
t = timed(lambda: sim.get_probability([True, False], [3, 17]))
emit({"op": "get_probability 30q, 2 qubits fixed (a quarter of the amplitudes contribute)", "s": t,
      "GBs_if_every_amplitude_were_read": 16 * amps / t / 1e9, "GBs_of_contributing_amplitudes": 4 * amps / t / 1e9})
t = timed(lambda: sim.emulate_math_addConstant(12345, [list(range(4, 16))], [29]))
emit({"op": "emulate_math_addConstant 12-bit register, 1 control, 30q", "s": t, "GBs_algorithmic(32B/amp)": 32 * amps / t / 1e9})
t = timed(lambda: sim.emulate_math_multiplyByConstantModN(7, 4087, [list(range(4, 16))], [29]))
emit({"op": "emulate_math_multiplyByConstantModN(7, 4087) 12-bit register, 1 control, 30q", "s": t,
      "GBs_algorithmic(32B/amp)": 32 * amps / t / 1e9})
t0 = time.perf_counter()
bits = sim.measure_qubits([5])
sim.synchronize()
emit({"op": "measure_qubits (1 qubit) 30q", "s": time.perf_counter() - t0, "bit": int(bits[0])})
t0 = time.perf_counter()
sim.collapse_wavefunction([9], [True])
sim.synchronize()
emit({"op": "collapse_wavefunction (1 qubit) 30q", "s": time.perf_counter() - t0})
del sim
if ref:
    r = ref.Simulator(5)
    nr = 24
    for q in range(nr):
        r.allocate_qubit(q)
    H = (np.array([[1, 1], [1, -1]]) / np.sqrt(2)).tolist()
    for q in range(nr):
        r.apply_controlled_gate(H, [q], [])
    r.run()
    t0 = time.perf_counter(); r.get_probability([True, False], [3, 17]); t1 = time.perf_counter()
    r.emulate_math_multiplyByConstantModN(7, 4087, [list(range(4, 16))], [nr - 1]); t2 = time.perf_counter()
    r.measure_qubits([5]); t3 = time.perf_counter()
    emit({"op": "reference C++ at 24 qubits (host)", "get_probability_s": t1 - t0, "emulate_math_mul_s": t2 - t1,
          "measure_qubits_s": t3 - t2, "scale_to_30q": 64})
