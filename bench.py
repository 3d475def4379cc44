#!/usr/bin/env python
"""Headline benchmark: BASELINE.json config 2 — random brickwork circuit (Rx/Ry/Rz + CNOT/CZ), complex128, gate fusion.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = the whole circuit (depth 20; 890 gates at 30 qubits) applied once to a resident 2^n state.
  N = 1 : n = 30 qubits (16 GiB state).  N > 1 (torchrun, one rank per GPU): weak scaling, n = 30 + log2 N, the state is
  sharded by log2 N qubits and dense gates on those qubits trigger global<->local remaps over NVLink.
Metric: amp-updates/s = gates * 2^n / t, whole job.

The JSON line carries
  value            device-timed (CUDA events on the engine's stream), state resident in HBM, packed gate stream handed over
                   in one C-ABI call per step;
  e2e              the same circuit gate by gate through the reference-facing native seam from host NumPy matrices, one
                   probability read back per step;
  e2e_main_engine  (N = 1) the same circuit as ProjectQ gates through MainEngine(projectq_b200.Simulator(gate_fusion=True),
                   engine_list=[]) — the drop-in number of SURVEY §8d; needs the reference's Python front end, staged under
                   baseline/_ref by oracle/Makefile;
  roofline         the dominant kernel's own launch time (every dense launch bracketed by CUDA events in a separate profiled
                   pass over the same steps) against the measured HBM peak;
  parity           computed outside the timed regions: U then U^dagger must restore the seeded initial state at sampled
                   indices, and the sampled amplitudes after one step must equal the committed 1-GPU fixture of the same
                   circuit (tests/golden/brickwork_1gpu_<n>q.json) — the sharded run is checked against a single-GPU run;
  strong_33q       BASELINE config 5: the 33-qubit circuit on N GPUs (128 GiB of state in total);
  weak_36q         (N = 8) the 36-qubit circuit, 1.1 TB of state: sec/layer.
`--impl reference` times the unmodified reference C++ simulator (oracle/_ref/_cppsim, OpenMP on all host cores) on the SAME
configuration (30 qubits when the host has the memory); one reference step is one brickwork layer of the circuit — a
bounded sample, the metric is gates * 2^n / t either way.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from projectq_b200.workloads import brickwork_circuit, inverse_circuit, pack_gate_stream  # noqa: E402

DEPTH = 20
BASE_QUBITS = 30
N_PARITY_SAMPLES = 256
STATE_SEED = 2026


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def host_memory_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 2**20
    except OSError:
        pass
    return 0.0


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference C++ simulator on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def load_reference_module():
    import glob
    import importlib.util

    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "_cppsim*.so"))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("_cppsim", hits[0])
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except ImportError:
        return None
    return mod


def reference_layers(n_qubits, steps, warmup, gate_fusion=True):
    """Time the reference on the config-2 circuit at n_qubits, one brickwork layer per step (layers taken in circuit order,
    cyclically).  Returns (amp-updates/s over the timed steps, ms per step, info)."""
    # the reference is an OpenMP code: give it every host core (torchrun exports OMP_NUM_THREADS=1 to its workers, which
    # would silently turn the baseline into a single-thread run); libgomp reads this when the module is first loaded
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.setdefault("OMP_PROC_BIND", "spread")  # the reference's own advice, _simulator.py:50-55
    mod = load_reference_module()
    if mod is None:
        return None, None, "oracle/_ref/_cppsim not built"
    gates = brickwork_circuit(n_qubits, DEPTH)
    # split into layers: a layer = n single-qubit rotations followed by its two-qubit gates
    layers, cur = [], []
    for m, t, c in gates:
        if not c and len(cur) >= n_qubits and cur[-1][2]:
            layers.append(cur)
            cur = []
        cur.append((m.tolist(), t, c))
    layers.append(cur)
    assert len(layers) == DEPTH and sum(len(x) for x in layers) == len(gates)
    sim = mod.Simulator(1)
    t0 = time.perf_counter()
    for q in range(n_qubits):
        sim.allocate_qubit(q)
    alloc_s = time.perf_counter() - t0
    times, n_gates = [], 0
    for it in range(warmup + steps):
        layer = layers[it % DEPTH]
        t0 = time.perf_counter()
        for m, t, c in layer:
            sim.apply_controlled_gate(m, t, c)
            if not gate_fusion:
                sim.run()
        sim.run()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            n_gates += len(layer)
    del sim
    value = n_gates * float(1 << n_qubits) / sum(times)
    info = {"cores": cores, "kind": "reference", "gate_fusion": gate_fusion,
            "sample": "config-2 generator at %d qubits, %d timed steps of one brickwork layer each (%d gates), after %d "
                      "warm-up layers; allocate_qubit x %d took %.1f s" % (n_qubits, steps, n_gates, warmup, n_qubits, alloc_s)}
    return value, 1e3 * sum(times) / len(times), info


def run_reference(args, rank):
    if rank != 0:
        return
    # the same configuration as our arm when the host can hold it (16 GiB state + the reference's 16 GiB grow buffer)
    n = 30 if host_memory_gb() >= 56 else 28
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # a 30-qubit layer takes the reference several seconds: keep the whole run within a few minutes whatever K and W are
    budget_layers = 30 if n == 30 else 100
    if steps + warmup > budget_layers:
        warmup = min(warmup, 2)
        steps = min(steps, budget_layers - warmup)
    value, ms, info = reference_layers(n, steps, warmup)
    if value is None:
        print(json.dumps({"impl": "reference", "unavailable": info}))
        return
    line = {
        "impl": "reference", "metric": "brickwork amp-updates/s", "value": value, "unit": "amp-updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "random brickwork circuit, %d qubits, depth %d (Rx/Ry/Rz + CNOT/CZ), complex128, gate_fusion=True; "
                               "one reference step = one layer of the circuit" % (n, DEPTH), "qubits": n},
        "cpu_baseline": dict(info, value=value, unit="amp-updates/s"),
        "e2e": {"value": value, "unit": "amp-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
class Plumbing:
    """torch.distributed (gloo) is used only to broadcast NCCL ids, to take the max of timings and as a barrier"""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("gloo", rank=self.rank, world_size=self.world)
            self.dist = dist

    def allmax(self, x):
        if self.dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def gather(self, obj):
        """list of every rank's obj (on every rank)"""
        if self.dist is None:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def make_sim(self, fusion=0):
        from projectq_b200.backend import SimulatorBackend, nccl_unique_id

        opts = dict(device=self.local_rank)
        if fusion:
            opts["fusion_max_qubits"] = fusion
        if self.world > 1:
            box = [nccl_unique_id() if self.rank == 0 else None]  # one fresh NCCL id per engine
            self.dist.broadcast_object_list(box, src=0)
            opts.update(rank=self.rank, world_size=self.world, nccl_unique_id=box[0])
        return SimulatorBackend(1, **opts)


def parity_block(sim, n, gates, body, n_gates):
    """outside every timed region.  (1) fixture: amplitudes after ONE step from the seeded state vs the committed single-GPU
    run of the same circuit; (2) U then U^dagger restores the seeded state.  Leaves the state = seeded state (+ rounding)."""
    rng = np.random.default_rng(n)
    idx = rng.integers(0, 1 << n, N_PARITY_SAMPLES, dtype=np.uint64)
    sim.init_random_state(n, STATE_SEED)
    before = np.asarray(sim.get_amplitudes(idx))
    sim.apply_gate_stream(body, n_gates, True)
    sim.run()
    after_u = np.asarray(sim.get_amplitudes(idx))
    out = {"samples": N_PARITY_SAMPLES}
    fixture = os.path.join(ROOT, "tests", "golden", "brickwork_1gpu_%dq.json" % n)
    if os.path.exists(fixture):
        with open(fixture) as f:
            fx = json.load(f)
        ref = np.array([complex(a, b) for a, b in fx["amplitudes"]])
        assert fx["indices"] == [int(i) for i in idx]
        out["vs_1gpu_fixture_max_abs"] = float(np.max(np.abs(after_u - ref)))
        out["fixture"] = os.path.relpath(fixture, ROOT)
    else:
        out["vs_1gpu_fixture_max_abs"] = None
    inv_body, inv_n = pack_gate_stream(inverse_circuit(gates))
    sim.apply_gate_stream(inv_body, inv_n, True)
    sim.run()
    back = np.asarray(sim.get_amplitudes(idx))
    out["u_udagger_max_abs"] = float(np.max(np.abs(back - before)))
    out["moved_by_u_max_abs"] = float(np.max(np.abs(after_u - before)))  # the circuit is not the identity
    out["norm_after"] = sim.norm_squared()
    worst = max(out["u_udagger_max_abs"], out["vs_1gpu_fixture_max_abs"] or 0.0)
    out["max_abs"] = worst
    out["ok"] = bool(worst < 1e-12 and out["moved_by_u_max_abs"] > 1e-7 and abs(out["norm_after"] - 1.0) < 1e-12)
    return out


def timed_steps(pl, sim, body, n_gates, steps, warmup, clocks=None):
    def step():
        sim.apply_gate_stream(body, n_gates, True)
        sim.run()  # every step is self-contained: nothing stays queued past the timed region

    for _ in range(warmup):
        step()
    sim.synchronize()
    pl.barrier()
    sim.reset_stats()
    if clocks is not None:
        clocks.start()
    sim.synchronize()
    pl.barrier()
    sim.timer_start()
    for _ in range(steps):
        step()
    ms_total = sim.timer_stop()
    pl.barrier()
    clock_info = clocks.stop() if clocks is not None else None
    st = sim.stats()
    # the ranks run in lock step (every remap is a rendezvous), so the time a rank spends waiting for an exchange is partly
    # the time it waits for a slower GPU; every rank's figures are reported, and the exposed part of the remaps is the
    # smallest wait over the ranks (the slowest GPU waits for nobody)
    st["per_rank"] = pl.gather({"remap_ms": st["remap_ms"], "remap_comm_ms": st["remap_comm_ms"], "step_ms": ms_total / steps})
    return pl.allmax(ms_total) / steps, st, clock_info


def remap_report(st, steps, ms_per_step):
    sent = st["remap_bytes_sent"] / steps
    per_rank = st["per_rank"]
    comm_ms = min(r["remap_comm_ms"] for r in per_rank) / steps  # shortest = least time spent waiting for peers to arrive
    exposed = min(r["remap_ms"] for r in per_rank) / steps
    return {"main_stream_wait_ms_per_step_by_rank": [r["remap_ms"] / steps for r in per_rank],
            "comm_ms_per_step_by_rank": [r["remap_comm_ms"] / steps for r in per_rank],
            "device_ms_per_step_by_rank": [r["step_ms"] for r in per_rank],"per_step": st["remaps"] / steps, "qubits_moved_per_step": st["remap_qubits"] / steps,
            "peer_memory_kernel": st["p2p_remaps"] / steps, "pipelined": st["pipelined_remaps"] / steps,
            "bytes_sent_per_gpu_per_step": sent, "comm_ms_per_step": comm_ms,
            "GBs_per_direction": sent / (comm_ms * 1e-3) / 1e9 if comm_ms > 0 else None,
            "exposed_ms_per_step": exposed, "exposed_frac_of_step": exposed / ms_per_step, "nvlink_peak_GBs": 900.0}


def extra_leg(pl, n, steps, warmup):
    """one more circuit size on a fresh engine (33-qubit strong-scaling point, 36-qubit weak-scaling point)"""
    gates = brickwork_circuit(n, DEPTH)
    body, n_gates = pack_gate_stream(gates)
    sim = pl.make_sim()
    parity = parity_block(sim, n, gates, body, n_gates)
    ms, st, _ = timed_steps(pl, sim, body, n_gates, steps, warmup)
    out = {"qubits": n, "gates": n_gates, "steps": steps, "warmup": warmup, "ms_per_step": ms,
           "value": n_gates * float(1 << n) / (ms * 1e-3), "unit": "amp-updates/s", "sec_per_layer": ms * 1e-3 / DEPTH,
           "state_GiB_per_gpu": 16.0 * float(1 << n) / pl.world / 2**30,
           "fused_passes_per_step": (sum(st["dense_passes"]) + st["diag_passes"]) / steps, "parity": parity}
    if pl.world > 1:
        out["remap"] = remap_report(st, steps, ms)
    del sim
    return out


def main_engine_leg(n, steps):
    """config 2 as ProjectQ gates through MainEngine(Simulator(gate_fusion=True), engine_list=[]) on the CUDA backend"""
    pkg = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(pkg, "projectq")):
        return {"unavailable": "the reference's Python front end is not staged under baseline/_ref (run `make -C oracle`)"}
    import types

    if "matplotlib" not in sys.modules:  # projectq's circuit drawer imports it at package import; not used here
        mpl = types.ModuleType("matplotlib")
        for sub, names in {"pyplot": [], "collections": ["LineCollection", "PatchCollection"], "lines": ["Line2D"],
                           "patches": ["Circle", "Arc", "Rectangle"]}.items():
            m = types.ModuleType("matplotlib." + sub)
            for nm in names:
                setattr(m, nm, type(nm, (), {}))
            setattr(mpl, sub, m)
            sys.modules["matplotlib." + sub] = m
        sys.modules["matplotlib"] = mpl
    sys.path.insert(0, pkg)
    try:
        from projectq import MainEngine
        from projectq.ops import CNOT, CZ, Rx, Ry, Rz

        from projectq_b200 import Simulator
    except ImportError as exc:
        return {"unavailable": "cannot import projectq: %s" % (exc,)}
    sim = Simulator(gate_fusion=True, rnd_seed=1)
    eng = MainEngine(sim, engine_list=[])
    q = eng.allocate_qureg(n)
    eng.flush()

    def step():
        rng = np.random.default_rng(2026)
        for d in range(DEPTH):
            for i in range(n):
                kind = int(rng.integers(0, 3))
                th = float(rng.uniform(0, 2 * np.pi))
                (Rx, Ry, Rz)[kind](th) | q[i]
            for i in range(d % 2, n - 1, 2):
                if int(rng.integers(0, 2)) == 0:
                    CNOT | (q[i], q[i + 1])
                else:
                    CZ | (q[i], q[i + 1])
        eng.flush()
        return sim.get_probability("0", [q[0]])

    step()
    sim._simulator.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        p0 = step()
    sim._simulator.synchronize()
    dt = (time.perf_counter() - t0) / steps
    n_gates = DEPTH * n + sum(len(range(d % 2, n - 1, 2)) for d in range(DEPTH))
    out = {"value": n_gates * float(1 << n) / dt, "unit": "amp-updates/s", "ms_per_step": dt * 1e3, "steps": steps, "p0": p0,
           "path": "MainEngine(projectq_b200.Simulator(gate_fusion=True), engine_list=[]) -> pybind shim -> C ABI -> CUDA"}
    # leave a measurable state behind so that MainEngine's shutdown (deallocation of every qubit) succeeds
    from projectq.ops import All, Measure

    All(Measure) | q
    eng.flush()
    return out


def run_ours(args):
    pl = Plumbing()
    rank, world = pl.rank, pl.world
    warmup = max(args.warmup, 3)
    n = args.qubits or (BASE_QUBITS + int(round(np.log2(max(world, 1)))))
    gates = brickwork_circuit(n, DEPTH)
    body, n_gates = pack_gate_stream(gates)
    amps_total = float(1 << n)
    sim = pl.make_sim(args.fusion)

    # ---- parity (outside the timed region) ----
    parity = parity_block(sim, n, gates, body, n_gates)

    # ---- headline: device-timed, state resident ----
    clocks = ClockSampler(pl.local_rank) if rank == 0 else None
    ms_per_step, st, clock_info = timed_steps(pl, sim, body, n_gates, args.steps, warmup, clocks)
    value = n_gates * amps_total / (ms_per_step * 1e-3)

    # ---- the dominant kernel's own launch time: same steps, every dense launch bracketed by events ----
    prof_steps = max(1, min(args.steps, 3))
    sim.set_profiling(True)
    _, st_prof, _ = timed_steps(pl, sim, body, n_gates, prof_steps, 1)
    sim.set_profiling(False)

    # ---- e2e: gate by gate through the backend API from host matrices, a probability read back every step ----
    host_gates = [(np.ascontiguousarray(m), t, c) for m, t, c in gates]
    h2d = sum(m.nbytes + 4 * (len(t) + len(c)) for m, t, c in host_gates)

    def step_api():
        for m, t, c in host_gates:
            sim.apply_controlled_gate(m, t, c)
        sim.run()
        return sim.get_probability([False], [0])

    step_api()
    sim.synchronize()
    pl.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        p0 = step_api()
    sim.synchronize()
    e2e_s = pl.allmax(time.perf_counter() - t0) / args.steps
    e2e_value = n_gates * amps_total / e2e_s
    norm = sim.norm_squared()
    del sim

    extras = {}
    if not args.no_extra_legs:
        legs = ([("strong_33q", 33)] if n != 33 else []) + ([("weak_36q", 36)] if world == 8 else [])
        for key, qubits in legs:
            try:
                extras[key] = extra_leg(pl, qubits, 3, 2)
            except (RuntimeError, MemoryError) as exc:  # e.g. not enough free HBM for a 128 GiB shard: the headline stands
                if world > 1:
                    raise  # the other ranks are inside collectives: fail loudly rather than hang
                extras[key] = {"unavailable": str(exc)[:300]}
    me = None
    if world == 1 and not args.no_main_engine:
        try:
            me = main_engine_leg(n, max(1, min(args.steps, 3)))
        except Exception as exc:  # the front end is the reference's code: report, do not lose the line
            me = {"unavailable": "%s: %s" % (type(exc).__name__, str(exc)[:300])}

    if rank != 0:
        return
    peaks, peak_kind = measured_peaks()
    passes = st["dense_passes"]
    n_pass = (sum(passes) + st["diag_passes"]) / args.steps
    dom_k = int(np.argmax(passes))
    local_amps = amps_total / world
    # dominant kernel: the dense k-qubit apply; every launch reads and writes each local amplitude once (32 B/amp)
    n_dom = st_prof["dense_passes"][dom_k]
    launch_ms = st_prof["pass_ms"][dom_k] / max(n_dom, 1)
    achieved = 32.0 * local_amps / (launch_ms * 1e-3) / 1e9
    dense_ms_per_step = (sum(st_prof["pass_ms"]) + st_prof["diag_ms"]) / prof_steps
    line = {
        "metric": "brickwork amp-updates/s", "value": value, "unit": "amp-updates/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "random brickwork circuit, %d qubits, depth %d (Rx/Ry/Rz + CNOT/CZ), complex128, fused "
                               "dense passes" % (n, DEPTH),
                   "qubits": n, "gates": n_gates, "fused_passes_per_step": n_pass, "pass_width_histogram": passes,
                   "diag_passes": st["diag_passes"], "l2": "state (%.1f GiB per GPU) is far larger than the 126 MB L2"
                   % (16.0 * local_amps / 2**30), "parallelism": "state sharded over %d rank bits" % int(np.log2(world)),
                   "norm_after": norm, "p0": p0, "sec_per_layer": ms_per_step * 1e-3 / DEPTH},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_kind": peak_kind,
                     "kernel": "apply_dense_kernel<k=%d>" % dom_k, "avg_launch_ms": launch_ms,
                     "launches_timed": int(n_dom), "algorithmic_bytes_per_launch": 32.0 * local_amps,
                     "all_pass_launches_ms_per_step": dense_ms_per_step,
                     "share_of_step": dense_ms_per_step / ms_per_step,
                     "traffic_note": "not measurable inside the run; the ncu --set full capture of this kernel is summarised "
                                     "under profiles/ (dram read+write = 0.998 x algorithmic)"},
        "e2e": {"value": e2e_value, "unit": "amp-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                "ms_per_step": e2e_s * 1e3},
        "parity": parity,
        "gpu_launches": int(st["kernel_launches"]),
        "clocks": clock_info,
    }
    if me is not None:
        line["e2e_main_engine"] = me
    if world > 1:
        line["remap"] = remap_report(st, args.steps, ms_per_step)
    line.update(extras)
    if n == 33:
        line["strong_33q"] = {"same_as_headline": True, "qubits": 33, "ms_per_step": ms_per_step, "value": value,
                              "sec_per_layer": ms_per_step * 1e-3 / DEPTH, "parity": parity}
    if "weak_36q" in extras:
        line["sec_per_layer_36q"] = extras["weak_36q"]["sec_per_layer"]
    if world == 1 and not args.no_cpu_baseline:
        # bounded sample of the same workload on the host cores: 28 qubits, 1 warm-up + 3 timed layers (~10-30 s)
        v, _, info = reference_layers(28, 3, 1)
        line["cpu_baseline"] = dict(info, value=v, unit="amp-updates/s") if v else {"unavailable": info}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=0, help="override the qubit count (default 30 + log2 gpus)")
    ap.add_argument("--fusion", type=int, default=0, help="max fused width (0 = engine default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the 33-qubit / 36-qubit legs")
    ap.add_argument("--no-main-engine", action="store_true", help="skip the MainEngine e2e leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
