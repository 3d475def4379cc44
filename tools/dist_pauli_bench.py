"""Config 4 on a sharded state: TFIM <H>, H psi and exp(-iHt) with one process per GPU (torchrun), so that the X terms on the
rank-bit qubits read the partner shards over NVLink.  Prints wall times from rank 0.  Not part of the product path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_pauli_bench.py 30
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist

    from projectq_b200.backend import SimulatorBackend, nccl_unique_id
    from projectq_b200.workloads import tfim_terms

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 29
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    sim = SimulatorBackend(1, device=local, rank=rank, world_size=world, nccl_unique_id=box[0])
    sim.init_random_state(n, 7)
    terms = tfim_terms(n)
    ids = list(range(n))

    def wall(fn, reps=1):
        sim.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        sim.synchronize()
        dist.barrier()
        return (time.perf_counter() - t0) * 1e3 / reps, r

    wall(lambda: sim.get_expectation_value(terms, ids))
    t_e, e0 = wall(lambda: sim.get_expectation_value(terms, ids), 3)
    cterms = [(t, complex(c)) for t, c in terms]
    wall(lambda: sim.apply_qubit_operator(cterms, ids))
    t_a, _ = wall(lambda: sim.apply_qubit_operator(cterms, ids), 3)
    sim2 = None
    del sim
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    sim = SimulatorBackend(1, device=local, rank=rank, world_size=world, nccl_unique_id=box[0])
    sim.init_random_state(n, 7)
    e_before = sim.get_expectation_value(terms, ids)
    wall(lambda: sim.emulate_time_evolution(terms, 0.02, ids, []))
    t_t, _ = wall(lambda: sim.emulate_time_evolution(terms, 0.1, ids, []))
    e_after = sim.get_expectation_value(terms, ids)
    if rank == 0:
        print(json.dumps({"config": "TFIM on a sharded state", "qubits": n, "gpus": world, "terms": len(terms), "expectation_ms": t_e,
                          "apply_ms": t_a, "time_evolution_t0.1_ms": t_t, "energy_before": e_before, "energy_after": e_after,
                          "energy_drift": abs(e_after - e_before), "stats": {k: v for k, v in sim.stats().items() if "remap" in k}}),
              flush=True)


if __name__ == "__main__":
    main()
