"""CPU (gloo) check of the N > 1 plumbing, run with world_size 2:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_plumbing_cpu.py
What runs here is the host side of the sharded path: the id broadcast bench.py/dist_check.py use, the max-over-ranks timing
reduction, and the SPMD determinism of fuser + remap planner (every rank must derive the same passes and the same swaps
from the same gate stream, or the NCCL exchanges would not pair up)."""
import ctypes
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import brickwork_circuit, pack_gate_stream  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # 1. id broadcast (a stand-in 128-byte id: creating a real one needs a GPU)
    box = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    assert box[0] == bytes(range(128))
    # 2. max-over-ranks timing reduction
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t[0]) == 10.0 + world - 1
    # 3. fuser + planner determinism across ranks
    lib = ctypes.CDLL(os.path.join(ROOT, "projectq_b200", "libpqb200.so"))
    n, g = 16, int(np.log2(world))
    gates = brickwork_circuit(n, 6, seed=12)
    body, cnt = pack_gate_stream(gates)
    cap = len(body) * 8 + (1 << 20)
    out = ctypes.create_string_buffer(cap)
    used, passes = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.pqb_host_fuse_stream(body, ctypes.c_size_t(len(body)), ctypes.c_size_t(cnt), 4, out, ctypes.c_size_t(cap),
                                    ctypes.byref(used), ctypes.byref(passes)) == 0
    raw = out.raw[: used.value]
    # walk the passes, planning a remap whenever a target sits on a rank bit (logical top g qubits start global)
    L = n - g
    loc = (ctypes.c_uint8 * n)(*[p if p < L else 64 + (p - L) for p in range(n)])
    swaps = []
    off = 0
    for _ in range(passes.value):
        k, nc = np.frombuffer(raw, dtype=np.uint32, count=2, offset=off)
        ids = np.frombuffer(raw, dtype=np.uint32, count=int(k + nc), offset=off + 8)
        off += 8 + 4 * int(k + nc) + 16 * (1 << int(k)) ** 2
        need = (ctypes.c_uint32 * int(k))(*[int(x) for x in ids[:k]])
        pairs = (ctypes.c_int32 * 16)()
        npairs = ctypes.c_size_t()
        assert lib.pqb_host_plan_remap(loc, ctypes.c_size_t(n), ctypes.c_int(L), need, ctypes.c_size_t(int(k)), pairs,
                                       ctypes.c_size_t(8), ctypes.byref(npairs)) == 0
        swaps += [int(pairs[i]) for i in range(2 * npairs.value)]
        assert all(loc[int(q)] < 64 for q in ids[:k])
    digest = hashlib.sha256(raw + bytes(np.array(swaps, dtype=np.int32).tobytes()) + bytes(loc)).hexdigest()
    gathered = [None] * world
    dist.all_gather_object(gathered, digest)
    assert len(set(gathered)) == 1, gathered
    assert len(swaps) > 0
    dist.barrier()
    if rank == 0:
        print("plumbing OK: %d passes, %d swaps" % (passes.value, len(swaps) // 2))


if __name__ == "__main__":
    main()
