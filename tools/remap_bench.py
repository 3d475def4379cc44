"""Remap micro-benchmark (run under torchrun, one rank per GPU): forces global<->local qubit remaps on a sharded state and
reports bytes sent per GPU / device time against the NVLink roofline.  Measurement tool, not product code."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist

    from projectq_b200.backend import SimulatorBackend, nccl_unique_id
    from tests.helpers import rand_unitary

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    n_local = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    low = len(sys.argv) > 2 and sys.argv[2] == "low"  # evict the lowest local bits (scattered halves -> packed exchange)
    g = int(np.log2(world))
    n = n_local + g
    dist.init_process_group("gloo", rank=rank, world_size=world)
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    sim = SimulatorBackend(1, device=local, rank=rank, world_size=world, nccl_unique_id=box[0])
    sim.init_random_state(n, 7)
    rng = np.random.default_rng(1)
    m = rand_unitary(rng, 1)
    # warm-up: one remap per global qubit (NCCL connection set-up happens here)
    for rep in range(2):
        for q in range(n - g, n):
            sim.apply_controlled_gate(m, [q], [])
            sim.run()
        for q in range(g):  # bring the original layout back: the evicted qubits were the top local ones
            sim.apply_controlled_gate(m, [n - g - 1 - q], [])
            sim.run()
    sim.synchronize()
    sim.reset_stats()
    reps = 3
    for rep in range(reps):
        # a k-qubit gate on all global qubits at once -> one multi-bit remap
        tq = list(range(n - g, n))
        sim.apply_controlled_gate(rand_unitary(rng, len(tq)), tq, [])
        if low:  # pending work on every local qubit except the lowest g: those become the eviction victims
            for q in range(g, n - g):
                sim.apply_controlled_gate(m, [q], [])
        sim.run()
        if low:
            sim.synchronize()
            st = sim.stats()
            if rank == 0:
                print(json.dumps({"mode": "low", "remaps": st["remaps"], "GBs_per_direction": st["remap_bytes_sent"] / max(st["remap_ms"], 1e-9) / 1e6}))
            continue
        tq = list(range(n - 2 * g, n - g))
        sim.apply_controlled_gate(rand_unitary(rng, len(tq)), tq, [])
        sim.run()
    sim.synchronize()
    st = sim.stats()
    nrm = sim.norm_squared()
    if rank == 0:
        gbs = st["remap_bytes_sent"] / max(st["remap_ms"], 1e-9) / 1e6
        print(json.dumps({"world": world, "qubits": n, "shard_GiB": 16.0 * (1 << n_local) / 2**30, "remaps": st["remaps"], "p2p_remaps": st.get("p2p_remaps"),
                          "bytes_sent_per_gpu": st["remap_bytes_sent"], "remap_ms": st["remap_ms"],
                          "GBs_per_direction": gbs, "of_nominal_900": gbs / 900.0, "of_measured_770": gbs / 770.0,
                          "norm": nrm, "nccl_env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}))
    dist.barrier()


if __name__ == "__main__":
    main()
