#!/usr/bin/env python
"""Headline benchmark: BASELINE.json config 2 — random brickwork circuit (Rx/Ry/Rz + CNOT/CZ), complex128, gate fusion.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = the whole circuit (depth 20; 890 gates at 30 qubits) applied once to a resident 2^n state.
  N = 1 : n = 30 qubits (16 GiB state).  N > 1 (torchrun, one rank per GPU): weak scaling, n = 30 + log2 N, the state is
  sharded by the top log2 N qubits and dense gates on those qubits trigger global<->local remaps over NVLink.
Metric: amp-updates/s = gates * 2^n / t, whole job.  `value` is device-timed (CUDA events on the engine's
stream) with the state resident in HBM and the packed gate stream handed over in one C-ABI call; `e2e` drives the same
circuit gate by gate through the reference-facing backend API from host NumPy matrices and reads a probability back.
`--impl reference` times the unmodified reference C++ simulator (oracle/_ref/_cppsim, OpenMP on all host cores) on a
bounded sample of the same workload (same generator, fewer qubits: the metric is per amplitude).
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

from tests.helpers import brickwork_circuit, pack_gate_stream  # noqa: E402

DEPTH = 20
BASE_QUBITS = 30
REF_SAMPLE_QUBITS = 24


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


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


def reference_sample(n_qubits, steps, warmup):
    """time the unmodified reference C++ simulator on the same generator at n_qubits; returns (best amp-updates/s, info)"""
    from tests.conftest import load_ref_cppsim

    # the reference is an OpenMP code: give it every host core (torchrun exports OMP_NUM_THREADS=1 to its workers, which
    # would silently turn the baseline into a single-thread run); libgomp reads this when the module is first loaded
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.setdefault("OMP_PROC_BIND", "spread")  # the reference's own advice, _simulator.py:50-55
    mod = load_ref_cppsim()
    if mod is None:
        return None, "oracle/_ref/_cppsim not built"
    gates = [(m.tolist(), t, c) for m, t, c in brickwork_circuit(n_qubits, DEPTH)]
    best = {}
    for fusion in (False, True):
        sim = mod.Simulator(1)
        for q in range(n_qubits):
            sim.allocate_qubit(q)
        times = []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            for m, t, c in gates:
                sim.apply_controlled_gate(m, t, c)
                if not fusion:
                    sim.run()
            sim.run()
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
        best[fusion] = len(gates) * float(1 << n_qubits) * len(times) / sum(times)
        del sim
    fusion = max(best, key=best.get)
    info = {"cores": cores, "kind": "reference", "gate_fusion": fusion,
            "sample": "same brickwork generator at %d qubits, depth %d (%d gates), %d timed steps; gate_fusion off/on = "
                      "%.3g / %.3g amp-updates/s" % (n_qubits, DEPTH, len(gates), steps, best[False], best[True])}
    return best[fusion], info


def run_reference(args, rank):
    if rank != 0:
        return
    # bounded: 1 warm-up + at most 2 timed circuit runs per fusion setting keep the arm within a few minutes on 8 cores
    steps = max(1, min(args.steps, 2))
    value, info = reference_sample(REF_SAMPLE_QUBITS, steps, 1)
    if value is None:
        print(json.dumps({"impl": "reference", "unavailable": info}))
        return
    n_gates = len(brickwork_circuit(REF_SAMPLE_QUBITS, DEPTH))
    ms = n_gates * float(1 << REF_SAMPLE_QUBITS) / value * 1e3
    line = {
        "impl": "reference", "metric": "brickwork amp-updates/s", "value": value, "unit": "amp-updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "random brickwork circuit, depth 20 (Rx/Ry/Rz + CNOT/CZ), complex128; reference timed on a "
                               "%d-qubit sample of the 30-qubit workload" % REF_SAMPLE_QUBITS},
        "cpu_baseline": dict(info, value=value, unit="amp-updates/s"),
        "e2e": {"value": value, "unit": "amp-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
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
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    warmup = max(args.warmup, 3)
    from projectq_b200.backend import SimulatorBackend, nccl_unique_id

    dist = None
    uid = None
    if world > 1:
        # plumbing only: broadcast the NCCL id and reduce timings over a CPU (gloo) group
        import torch
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("gloo", rank=rank, world_size=world)
        box = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]

    def allmax(x):
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def barrier():
        if dist is not None:
            dist.barrier()

    n = args.qubits or (BASE_QUBITS + int(round(np.log2(max(world, 1)))))
    gates = brickwork_circuit(n, DEPTH)
    body, n_gates = pack_gate_stream(gates)
    opts = dict(device=local_rank)
    if args.fusion:
        opts["fusion_max_qubits"] = args.fusion
    if world > 1:
        opts.update(rank=rank, world_size=world, nccl_unique_id=uid)
    sim = SimulatorBackend(1, **opts)
    sim.init_random_state(n, 2026)
    amps_total = float(1 << n)

    def step_packed():
        sim.apply_gate_stream(body, n_gates, True)
        sim.run()  # every step is self-contained: nothing stays queued past the timed region

    for _ in range(warmup):
        step_packed()
    sim.synchronize()
    barrier()
    sim.reset_stats()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    sim.synchronize()
    barrier()
    sim.timer_start()
    for _ in range(args.steps):
        step_packed()
    ms_total = sim.timer_stop()
    barrier()
    clock_info = clocks.stop() if rank == 0 else None
    ms_total = allmax(ms_total)
    st = sim.stats()
    ms_per_step = ms_total / args.steps
    value = n_gates * amps_total / (ms_per_step * 1e-3)

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
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        p0 = step_api()
    sim.synchronize()
    e2e_s = allmax(time.perf_counter() - t0) / args.steps
    e2e_value = n_gates * amps_total / e2e_s
    norm = sim.norm_squared()

    if rank != 0:
        return
    peaks, peak_kind = measured_peaks()
    passes = st["dense_passes"]
    n_pass = (sum(passes) + st["diag_passes"]) / args.steps
    dom_k = int(np.argmax(passes))
    local_amps = amps_total / world
    # dominant kernel: the dense k-qubit apply; every launch reads and writes each local amplitude once (32 B/amp)
    remap_ms = st["remap_ms"] / args.steps
    launch_ms = (ms_per_step - remap_ms) / max(n_pass, 1)
    achieved = 32.0 * local_amps / (launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_traffic_k4_30q.json")
    if world == 1 and n == 30 and dom_k == 4 and os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)["traffic_bytes_per_launch"]  # dram read+write of one launch, ncu --set full capture
    line = {
        "metric": "brickwork amp-updates/s", "value": value, "unit": "amp-updates/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "random brickwork circuit, %d qubits, depth %d (Rx/Ry/Rz + CNOT/CZ), complex128, fused "
                               "dense passes" % (n, DEPTH),
                   "qubits": n, "gates": n_gates, "fused_passes_per_step": n_pass, "pass_width_histogram": passes,
                   "diag_passes": st["diag_passes"], "l2": "state (%.1f GiB per GPU) is far larger than the 126 MB L2"
                   % (16.0 * local_amps / 2**30), "parallelism": "state sharded by the top %d qubits" % int(np.log2(world)),
                   "norm_after": norm, "p0": p0, "sec_per_layer": ms_per_step * 1e-3 / DEPTH},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_kind": peak_kind,
                     "kernel": "apply_dense_kernel<k=%d>" % dom_k, "avg_launch_ms": launch_ms,
                     "algorithmic_bytes_per_launch": 32.0 * local_amps},
        "e2e": {"value": e2e_value, "unit": "amp-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                "ms_per_step": e2e_s * 1e3},
        "gpu_launches": int(st["kernel_launches"]),
        "clocks": clock_info,
    }
    if world > 1:
        line["remap"] = {"per_step": st["remaps"] / args.steps, "bytes_sent_per_gpu_per_step": st["remap_bytes_sent"] / args.steps,
                         "ms_per_step": remap_ms,
                         "GBs_per_direction": (st["remap_bytes_sent"] / max(st["remap_ms"], 1e-9)) / 1e6 if st["remaps"] else None,
                         "nvlink_peak_GBs": 900.0}
    if world == 1 and not args.no_cpu_baseline:
        v, info = reference_sample(REF_SAMPLE_QUBITS, 1, 1)
        line["cpu_baseline"] = dict(info, value=v, unit="amp-updates/s") if v else {"unavailable": info}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
