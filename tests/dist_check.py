"""Sharded-state parity check, run as one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Every rank drives the same call sequence (SPMD) on its shard; results are compared with the NumPy oracle computed on every
rank.  torch.distributed (gloo) is plumbing only: it broadcasts the NCCL id.  Exit code 0 = all checks passed on all ranks.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.statevec_oracle import OracleSimulator  # noqa: E402
from tests.helpers import brickwork_circuit, rand_state, rand_unitary, tfim_terms  # noqa: E402

TOL = 1e-12


def main():
    import torch.distributed as dist

    from projectq_b200.backend import SimulatorBackend, nccl_unique_id

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def make(seed=1, **kw):
        box = [nccl_unique_id() if rank == 0 else None]  # one fresh NCCL id per communicator
        dist.broadcast_object_list(box, src=0)
        return SimulatorBackend(seed, device=local, rank=rank, world_size=world, nccl_unique_id=box[0], **kw)

    def same(gpu, chk, what, tol=TOL):
        m1, v1 = gpu.cheat()
        m2, v2 = chk.cheat()
        assert dict(m1) == dict(m2), (what, m1, m2)
        err = float(np.max(np.abs(np.asarray(v1) - v2)))
        assert err < tol, (what, err)
        return err

    rng = np.random.default_rng(99)  # same stream on every rank
    expect = os.environ.get("PQB_EXPECT", "pipelined")  # which remap path this run must have taken

    # 0. a lazily allocating program: gate on the first qubit right after its allocation, one run() per gate (the default
    #    gate_fusion=False path) — the first qubits must land on local bits, not on rank bits with nothing to evict
    gpu, chk = make(5), OracleSimulator(5)
    for q in range(4):
        gpu.allocate_qubit(q)
        chk.allocate_qubit(q)
        for m, t, c in ((rand_unitary(rng, 1), [q], []), (rand_unitary(rng, 1), [q], [q - 1] if q else [])):
            gpu.apply_controlled_gate(m, t, c)
            gpu.run()
            chk.apply_controlled_gate(m, t, c)
    m = rand_unitary(rng, 4)
    gpu.apply_controlled_gate(m, [3, 1, 0, 2], [])
    chk.apply_controlled_gate(m, [3, 1, 0, 2], [])
    same(gpu, chk, "lazy allocation")
    assert list(gpu.measure_qubits([0, 1, 2, 3])) == list(chk.measure_qubits([0, 1, 2, 3]))
    del gpu

    # 1. allocate one by one (the first log2(world) qubits land on rank bits), gates on every qubit -> remaps
    n = 12
    gpu, chk = make(3), OracleSimulator(3)
    for q in range(n):
        gpu.allocate_qubit(q)
        chk.allocate_qubit(q)
        m = rand_unitary(rng, 1)
        gpu.apply_controlled_gate(m, [q], [q - 1] if q else [])
        chk.apply_controlled_gate(m, [q], [q - 1] if q else [])
    same(gpu, chk, "allocate+gates")
    for g in range(80):
        k = int(rng.integers(1, 5))
        nc = int(rng.integers(0, 3))
        qs = [int(x) for x in rng.permutation(n)[: k + nc]]
        m = rand_unitary(rng, k) if rng.random() < 0.8 else np.diag(np.exp(1j * rng.uniform(0, 6.28, 1 << k)))
        gpu.apply_controlled_gate(m, qs[:k], qs[k:])
        chk.apply_controlled_gate(m, qs[:k], qs[k:])
        if g % 9 == 0:
            gpu.run()
    same(gpu, chk, "random circuit")
    st = gpu.stats()
    assert st["remaps"] > 0, st
    if expect == "pipelined":
        assert st["pipelined_remaps"] > 0 and st["p2p_remaps"] > 0, st
    elif expect == "p2p":
        assert st["pipelined_remaps"] == 0 and st["p2p_remaps"] > 0, st
    elif expect == "nccl":
        assert st["p2p_remaps"] == 0, st

    # 1b. a longer random circuit with many controls (controls on rank-bit qubits switch whole ranks off, which must not
    #     change what the ranks of an exchange group agree on), flushed at random points; a second engine so that the first
    #     one keeps its state for the checks below
    n2 = 14
    gpu2, chk2 = make(7), OracleSimulator(7)
    for q in range(n2):
        gpu2.allocate_qubit(q)
        chk2.allocate_qubit(q)
    wf2 = rand_state(rng, n2)
    gpu2.set_wavefunction(wf2, list(range(n2)))
    chk2.set_wavefunction(wf2, list(range(n2)))
    for g in range(220):
        k = int(rng.integers(1, 4))
        nc = int(rng.integers(0, 4))
        qs = [int(x) for x in rng.permutation(n2)[: k + nc]]
        m = rand_unitary(rng, k) if rng.random() < 0.85 else np.diag(np.exp(1j * rng.uniform(0, 6.28, 1 << k)))
        gpu2.apply_controlled_gate(m, qs[:k], qs[k:])
        chk2.apply_controlled_gate(m, qs[:k], qs[k:])
        if rng.random() < 0.04:
            gpu2.run()
    same(gpu2, chk2, "random circuit with controls")
    del gpu2, chk2

    # 2. queries
    for _ in range(6):
        k = int(rng.integers(1, 5))
        sub = [int(x) for x in rng.permutation(n)[:k]]
        bits = [bool(b) for b in rng.integers(0, 2, k)]
        assert abs(gpu.get_probability(bits, sub) - chk.get_probability(bits, sub)) < TOL
        full = [int(x) for x in rng.permutation(n)]
        fb = [bool(b) for b in rng.integers(0, 2, n)]
        assert abs(gpu.get_amplitude(fb, full) - chk.get_amplitude(fb, full)) < TOL
    terms = tfim_terms(n) + [([(0, "Y"), (3, "X"), (11, "Z")], 0.37), ([], 0.25), ([(1, "Y"), (10, "Y")], -1.1)]
    ids = [int(x) for x in rng.permutation(n)]
    e1, e2 = gpu.get_expectation_value(terms, ids), chk.get_expectation_value(terms, ids)
    assert abs(e1 - e2) < 1e-11, (e1, e2)

    # 3. time evolution (controlled) and apply_qubit_operator
    # X/Y support of the operator must fit on one shard (n - log2(world) qubits); Z terms may touch every qubit
    nl = n - int(np.log2(world))
    hterms = tfim_terms(nl - 1) + [([(i, "Z"), (n - 2, "Z")], 0.3) for i in range(3)] + [([], 0.2)]
    gpu.emulate_time_evolution(hterms, 0.4, list(range(n - 1)), [n - 1])
    chk.emulate_time_evolution(hterms, 0.4, list(range(n - 1)), [n - 1])
    same(gpu, chk, "time evolution")
    terms = hterms + [([(0, "Y"), (3, "X"), (11, "Z")], 0.37), ([(1, "Y"), (2, "Y")], -1.1)]
    cterms = [(t, c * (1 + 0.5j)) for t, c in terms]
    gpu.apply_qubit_operator(cterms, ids)
    chk.apply_qubit_operator(cterms, ids)
    same(gpu, chk, "apply_qubit_operator", tol=1e-10)
    # 3b. operators that flip MORE qubits than one shard holds (a transverse field on every qubit): no remap can bring the
    #     whole X/Y support on-device, the partner amplitudes are read from the peers' shards over NVLink
    nrm = float(np.linalg.norm(chk.cheat()[1]))
    wfn = chk.cheat()[1] / nrm
    back0 = {v: k for k, v in chk.cheat()[0].items()}
    order0 = [back0[p] for p in range(n)]
    gpu.set_wavefunction(wfn, order0)
    chk.set_wavefunction(wfn, order0)
    full = tfim_terms(n) + [([(q, "Y" if q % 3 == 0 else "X") for q in range(n)], 0.21), ([(0, "Y"), (n - 1, "Z"), (5, "X")], -0.4)]
    if expect != "nccl":  # needs peer-mapped shards
        gpu.emulate_time_evolution(full, 0.05, list(range(n)), [])
        chk.emulate_time_evolution(full, 0.05, list(range(n)), [])
        same(gpu, chk, "time evolution, X on every qubit")
        cfull = [(t, c * (0.8 - 0.3j)) for t, c in full]
        gpu.apply_qubit_operator(cfull, ids)
        chk.apply_qubit_operator(cfull, ids)
        same(gpu, chk, "apply_qubit_operator, X on every qubit", tol=1e-10)
        e1, e2 = gpu.get_expectation_value(full, ids), chk.get_expectation_value(full, ids)
        assert abs(e1 - e2) < 1e-10 * max(1.0, abs(e2)), (e1, e2)
    else:
        try:
            gpu.get_expectation_value(full, ids)
            raise AssertionError("a string on every qubit needs peer-mapped shards: expected RuntimeError without them")
        except RuntimeError:
            pass
    nrm = float(np.linalg.norm(chk.cheat()[1]))
    wf = chk.cheat()[1] / nrm
    order = [int(x) for x in rng.permutation(n)]
    back = {v: k for k, v in chk.cheat()[0].items()}
    cur_order = [back[p] for p in range(n)]
    gpu.set_wavefunction(wf, cur_order)
    chk.set_wavefunction(wf, cur_order)
    same(gpu, chk, "set_wavefunction")
    del order

    # 4. emulate_math (bit-exact on identical inputs)
    gpu.emulate_math_addConstant(5, [[0, 1, 2, 3, 4]], [11])
    chk.emulate_math_addConstant(5, [[0, 1, 2, 3, 4]], [11])
    gpu.emulate_math_multiplyByConstantModN(3, 32, [[5, 6, 7, 8, 9]], [])
    chk.emulate_math_multiplyByConstantModN(3, 32, [[5, 6, 7, 8, 9]], [])
    m1, v1 = gpu.cheat()
    assert np.array_equal(np.asarray(v1), chk.cheat()[1]), "emulate_math not bit-exact"

    # 5. measurement (same RNG stream), collapse, deallocation incl. qubits on rank bits
    for mids in ([3], [0, 5], [7, 1, 2]):
        a, b = list(gpu.measure_qubits(mids)), list(chk.measure_qubits(mids))
        assert a == b, (mids, a, b)
    same(gpu, chk, "measure")
    for q in (3, 0, 5, 2, 1, 7):
        assert gpu.is_classical(q) == chk.is_classical(q)
        gpu.deallocate_qubit(q)
        chk.deallocate_qubit(q)
        same(gpu, chk, "deallocate %d" % q)
    gpu.collapse_wavefunction([4, 6], [True, False]) if chk.get_probability([True, False], [4, 6]) > 1e-6 else None
    if chk.get_probability([True, False], [4, 6]) > 1e-6:
        chk.collapse_wavefunction([4, 6], [True, False])
    same(gpu, chk, "collapse")
    gpu.allocate_qubit(50)
    chk.allocate_qubit(50)
    m = rand_unitary(rng, 2)
    gpu.apply_controlled_gate(m, [50, 4], [])
    chk.apply_controlled_gate(m, [50, 4], [])
    same(gpu, chk, "re-allocate")
    del gpu

    # 6. a wider state: fused brickwork on 20 qubits vs the oracle (several flushes: the qubits that left come back)
    n = 20
    gates = brickwork_circuit(n, 6, seed=8)
    gpu, chk = make(1), OracleSimulator(1)
    wf = rand_state(rng, n)
    for q in range(n):
        gpu.allocate_qubit(q)
        chk.allocate_qubit(q)
    gpu.set_wavefunction(wf, list(range(n)))
    chk.set_wavefunction(wf, list(range(n)))
    for i, (m, t, c) in enumerate(gates):
        gpu.apply_controlled_gate(m, t, c)
        chk.apply_controlled_gate(m, t, c)
        if i % 97 == 96:
            gpu.run()
    err = same(gpu, chk, "brickwork 20q")
    st = gpu.stats()

    # 7. checkpoint of the sharded state: every rank writes and reads its own shard file, nothing is gathered
    prefix = "/tmp/pqb_ckpt_%s" % os.environ.get("MASTER_PORT", "0")
    gpu.save_state(prefix)
    dist.barrier()
    other = make(77)
    other.load_state(prefix)
    same(other, chk, "checkpoint restore")
    m = rand_unitary(rng, 3)
    other.apply_controlled_gate(m, [0, 19, 7], [3])
    chk.apply_controlled_gate(m, [0, 19, 7], [3])
    same(other, chk, "gates after restore")
    assert list(other.measure_qubits([19, 2])) == list(chk.measure_qubits([19, 2]))
    del other
    dist.barrier()
    try:
        os.remove("%s.rank%dof%d.pqbs" % (prefix, rank, world))
    except OSError:
        pass
    if rank == 0:
        print("dist_check OK: world=%d, brickwork max|dpsi|=%.2e, stats=%s" % (world, err, st), flush=True)
    dist.barrier()


if __name__ == "__main__":
    main()
