"""CPU tests of the host-side pieces of the path: fuser, RNG replay, remap planner, C-ABI exports."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle.statevec_oracle import MT19937, OracleSimulator
from tests.helpers import brickwork_circuit, pack_gate_stream, rand_state, rand_unitary

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    path = os.path.join(ROOT, "projectq_b200", "libpqb200.so")
    if not os.path.exists(path):
        import __graft_entry__

        __graft_entry__.build()
    return ctypes.CDLL(path)


def test_c_abi_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "pqb200.h")).read()
    names = re.findall(r"PQB_API\s+[\w\s\*]+?\b(pqb_\w+)\s*\(", header)
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), name


def test_no_cpu_fallback_without_device(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    sim = ctypes.c_void_p()
    st = lib.pqb_create(ctypes.c_uint32(1), None, ctypes.byref(sim))
    assert st == 3  # PQB_ERR_CUDA
    lib.pqb_last_error.restype = ctypes.c_char_p
    assert b"no CUDA device" in lib.pqb_last_error(None)


def test_rng_stream_matches_libstdcxx_golden(lib):
    out = (ctypes.c_double * 6)()
    assert lib.pqb_host_rng_stream(ctypes.c_uint32(1), ctypes.c_size_t(6), out) == 0
    want = [0.99718480823026556, 0.93255736136816547, 0.128124447772306, 0.99904051546527362]
    assert list(out)[:4] == want
    r = MT19937(1)
    assert list(out) == [r.uniform01() for _ in range(6)]
    for seed in (0, 7, 4294967295):
        assert lib.pqb_host_rng_stream(ctypes.c_uint32(seed), ctypes.c_size_t(6), out) == 0
        r = MT19937(seed)
        assert list(out) == [r.uniform01() for _ in range(6)]


def fuse(lib, gates, max_qubits):
    body, n = pack_gate_stream(gates)
    cap = len(body) * 8 + (1 << 20)
    out = ctypes.create_string_buffer(cap)
    used = ctypes.c_size_t()
    passes = ctypes.c_size_t()
    st = lib.pqb_host_fuse_stream(body, ctypes.c_size_t(len(body)), ctypes.c_size_t(n), ctypes.c_int(max_qubits), out,
                                  ctypes.c_size_t(cap), ctypes.byref(used), ctypes.byref(passes))
    assert st == 0
    raw = out.raw[: used.value]
    res = []
    off = 0
    for _ in range(passes.value):
        k, nc = np.frombuffer(raw, dtype=np.uint32, count=2, offset=off)
        off += 8
        ids = np.frombuffer(raw, dtype=np.uint32, count=int(k + nc), offset=off)
        off += 4 * int(k + nc)
        d = 1 << int(k)
        m = np.frombuffer(raw, dtype=np.complex128, count=d * d, offset=off).reshape(d, d)
        off += 16 * d * d
        res.append((m, [int(x) for x in ids[:k]], [int(x) for x in ids[k:]]))
    assert off == used.value
    return res


def run_oracle(n, wf, gates):
    s = OracleSimulator(1)
    for q in range(n):
        s.allocate_qubit(q)
    s.set_wavefunction(wf, list(range(n)))
    for m, t, c in gates:
        s.apply_controlled_gate(m, t, c)
    return s.cheat()[1]


@pytest.mark.parametrize("max_qubits", [1, 2, 3, 4, 5])
def test_fuser_is_amplitude_equivalent_random(lib, max_qubits):
    rng = np.random.default_rng(40 + max_qubits)
    n = 9
    for trial in range(6):
        gates = []
        for g in range(50):
            k = int(rng.integers(1, max_qubits + 1))
            nc = int(rng.integers(0, 3))
            qs = [int(x) for x in rng.permutation(n)[: k + nc]]
            if rng.random() < 0.3:
                m = np.diag(np.exp(1j * rng.uniform(0, 6.28, 1 << k)))
            else:
                m = rand_unitary(rng, k)
            gates.append((m, qs[:k], qs[k:]))
        passes = fuse(lib, gates, max_qubits)
        assert all(len(t) <= max_qubits for _, t, _ in passes)
        wf = rand_state(rng, n)
        a = run_oracle(n, wf, gates)
        b = run_oracle(n, wf, passes)
        assert np.max(np.abs(a - b)) < 1e-12


def test_fuser_packs_brickwork_densely(lib):
    n, depth = 16, 20
    gates = brickwork_circuit(n, depth)
    passes = fuse(lib, gates, 5)
    assert len(passes) * 6 < len(gates)  # at least 6 gates per pass on average
    wf = rand_state(np.random.default_rng(1), n)
    assert np.max(np.abs(run_oracle(n, wf, gates) - run_oracle(n, wf, passes))) < 1e-12
    # width 1 = no fusion: one pass per gate, controls stay controls (CNOT = k=1 + control mask)
    p1 = fuse(lib, gates, 1)
    assert len(gates) - 2 * depth <= len(p1) <= len(gates)  # only back-to-back gates on one idle qubit merge
    assert sum(len(c) for _, _, c in p1) == sum(len(c) for _, _, c in gates)
    assert all(len(t) == 1 for _, t, _ in p1)


def test_fuser_keeps_common_controls_global(lib):
    rng = np.random.default_rng(9)
    gates = [(rand_unitary(rng, 1), [q], [7, 8]) for q in range(4)]
    passes = fuse(lib, gates, 5)
    assert len(passes) == 1
    m, t, c = passes[0]
    assert sorted(c) == [7, 8] and sorted(t) == [0, 1, 2, 3]
    # a gate lacking control 8 demotes it into the matrix
    gates.append((rand_unitary(rng, 1), [0], [7]))
    passes = fuse(lib, gates, 5)
    assert len(passes) == 1 and passes[0][2] == [7] and sorted(passes[0][1]) == [0, 1, 2, 3, 8]
    n = 9
    wf = rand_state(rng, n)
    assert np.max(np.abs(run_oracle(n, wf, gates) - run_oracle(n, wf, passes))) < 1e-12


def test_remap_planner(lib):
    # 2 rank bits, 6 local bits: logical 0,1 on rank bits 0,1; logical 2..7 on local bits 0..5
    loc = (ctypes.c_uint8 * 8)(64, 65, 0, 1, 2, 3, 4, 5)
    need = (ctypes.c_uint32 * 3)(0, 7, 3)
    pairs = (ctypes.c_int32 * 8)()
    n_pairs = ctypes.c_size_t()
    st = lib.pqb_host_plan_remap(loc, ctypes.c_size_t(8), ctypes.c_int(6), need, ctypes.c_size_t(3), pairs,
                                 ctypes.c_size_t(4), ctypes.byref(n_pairs))
    assert st == 0 and n_pairs.value == 1
    # logical 7 sits on the top local bit (5) and is needed, so the victim is local bit 4 (logical 6)
    assert (pairs[0], pairs[1]) == (0, 4)
    assert list(loc) == [4, 65, 0, 1, 2, 3, 64, 5]
    # everything local already -> no swaps
    st = lib.pqb_host_plan_remap(loc, ctypes.c_size_t(8), ctypes.c_int(6), need, ctypes.c_size_t(3), pairs,
                                 ctypes.c_size_t(4), ctypes.byref(n_pairs))
    assert st == 0 and n_pairs.value == 0
    # impossible: more needed qubits than local bits
    loc2 = (ctypes.c_uint8 * 3)(64, 0, 1)
    need2 = (ctypes.c_uint32 * 3)(0, 1, 2)
    st = lib.pqb_host_plan_remap(loc2, ctypes.c_size_t(3), ctypes.c_int(2), need2, ctypes.c_size_t(3), pairs,
                                 ctypes.c_size_t(4), ctypes.byref(n_pairs))
    assert st == 1


def test_exchange_plan_is_the_bit_swap_permutation(lib):
    """simulate the multi-bit remap on NumPy shards: following plan_exchange on every rank must realise exactly the
    permutation 'exchange rank bit r_i with local bit b_i' of the global index"""
    g_total, L = 3, 6
    world = 1 << g_total
    for swaps in ([(0, 5)], [(2, 0)], [(0, 5), (2, 1)], [(1, 0), (0, 3), (2, 4)]):
        shards = [np.arange(1 << L, dtype=np.int64) + (r << L) for r in range(world)]  # value = old global index
        new = [s.copy() for s in shards]
        pos = sorted(b for _, b in swaps)

        def members(pattern):
            idx = []
            for j in range(1 << (L - len(pos))):
                x = j
                for p in pos:
                    x = ((x >> p) << (p + 1)) | (x & ((1 << p) - 1))
                idx.append(x | pattern)
            return np.array(idx)

        pairs = (ctypes.c_int32 * (2 * len(swaps)))(*[v for rb in swaps for v in rb])
        plans = []
        for rank in range(world):
            peers = (ctypes.c_int32 * 8)()
            pats = (ctypes.c_uint64 * 8)()
            n = ctypes.c_size_t()
            assert lib.pqb_host_plan_exchange(rank, pairs, ctypes.c_size_t(len(swaps)), peers, pats, ctypes.c_size_t(8),
                                              ctypes.byref(n)) == 0
            assert n.value == (1 << len(swaps)) - 1
            plans.append({int(peers[i]): int(pats[i]) for i in range(n.value)})
        for rank in range(world):
            for peer, pattern in plans[rank].items():
                # this rank packs its sub-block `pattern` for the partner; the partner unpacks it into the sub-block it
                # exchanges with this rank (each rank sends from and receives into the same addresses)
                new[peer][members(plans[peer][rank])] = shards[rank][members(pattern)]
        for rank in range(world):
            for local in range(1 << L):
                old = int(new[rank][local])
                orank, olocal = old >> L, old & ((1 << L) - 1)
                # exchanging the bit pairs of (orank, olocal) must give (rank, local)
                er, el = orank, olocal
                for r, b in swaps:
                    rb, lb = (er >> r) & 1, (el >> b) & 1
                    er = (er & ~(1 << r)) | (lb << r)
                    el = (el & ~(1 << b)) | (rb << b)
                assert (er, el) == (rank, local)


def shard_schedule(lib, gates, n, rank_bits, flushes, max_qubits=4):
    body, cnt = pack_gate_stream(gates)
    cap = (len(body) * 8 + (1 << 20)) * flushes
    out = ctypes.create_string_buffer(cap)
    used = ctypes.c_size_t()
    st = lib.pqb_host_shard_schedule(body, ctypes.c_size_t(len(body)), ctypes.c_size_t(cnt), ctypes.c_uint32(n),
                                     ctypes.c_uint32(rank_bits), ctypes.c_int(max_qubits), ctypes.c_uint32(flushes), out,
                                     ctypes.c_size_t(cap), ctypes.byref(used))
    lib.pqb_last_error.restype = ctypes.c_char_p
    assert st == 0, lib.pqb_last_error(None)
    raw = out.raw[: used.value]
    events, off = [], 0
    while off < len(raw):
        k, nc = (int(x) for x in np.frombuffer(raw, dtype=np.uint32, count=2, offset=off))
        off += 8
        if k == 0xFFFFFFFF:
            pairs = np.frombuffer(raw, dtype=np.uint32, count=2 * nc, offset=off).reshape(nc, 2)
            off += 8 * nc
            events.append(("remap", [(int(a), int(b)) for a, b in pairs]))
        elif k == 0xFFFFFFFE:
            events.append(("flush", None))
        else:
            ids = np.frombuffer(raw, dtype=np.uint32, count=k + nc, offset=off)
            off += 4 * (k + nc)
            d = 1 << k
            m = np.frombuffer(raw, dtype=np.complex128, count=d * d, offset=off).reshape(d, d)
            off += 16 * d * d
            events.append(("pass", (m, [int(x) for x in ids[:k]], [int(x) for x in ids[k:]])))
    return events


@pytest.mark.parametrize("rank_bits", [1, 2, 3])
def test_sharded_schedule_is_equivalent_and_remaps_once_per_global_qubit(lib, rank_bits):
    """dry run of Engine::run_sharded: deferring the gates on off-device qubits must not change the circuit, every
    scheduled pass must act on on-device qubits only, and a brickwork flush needs one exchange per global qubit"""
    n, depth, flushes = 16, 6, 4  # depth well below n: one sweep of the chain per flush is possible
    gates = brickwork_circuit(n, depth, seed=21)
    events = shard_schedule(lib, gates, n, rank_bits, flushes)
    off_device = set(range(n - rank_bits, n))
    passes, remaps_per_flush, cur = [], [], 0
    for kind, payload in events:
        if kind == "pass":
            m, t, c = payload
            assert not (set(t) & off_device), "a dense target is off-device"
            passes.append(payload)
        elif kind == "remap":
            for incoming, evicted in payload:
                assert incoming in off_device and evicted not in off_device
                off_device.remove(incoming)
                off_device.add(evicted)
            cur += len(payload)
        else:
            remaps_per_flush.append(cur)
            cur = 0
    assert len(remaps_per_flush) == flushes
    assert all(r <= rank_bits for r in remaps_per_flush[1:]), remaps_per_flush  # steady state: one per global qubit
    assert remaps_per_flush[0] <= 2 * rank_bits
    wf = rand_state(np.random.default_rng(3), n)
    want = run_oracle(n, wf, gates * flushes)
    got = run_oracle(n, wf, passes)
    assert np.max(np.abs(want - got)) < 1e-12


def test_sharded_schedule_random_circuit(lib):
    rng = np.random.default_rng(77)
    n = 10
    gates = []
    for g in range(120):
        k = int(rng.integers(1, 4))
        nc = int(rng.integers(0, 3))
        qs = [int(x) for x in rng.permutation(n)[: k + nc]]
        gates.append((rand_unitary(rng, k), qs[:k], qs[k:]))
    events = shard_schedule(lib, gates, n, 3, 2, max_qubits=5)
    passes = [p for kind, p in events if kind == "pass"]
    wf = rand_state(rng, n)
    assert np.max(np.abs(run_oracle(n, wf, gates * 2) - run_oracle(n, wf, passes))) < 1e-12


def _fdpass_worker(rank, world, tag, q):
    lib = ctypes.CDLL(os.path.join(ROOT, "projectq_b200", "libpqb200.so"))
    q.put((rank, lib.pqb_host_fdpass_selftest(ctypes.c_uint64(tag), ctypes.c_int(rank), ctypes.c_int(world))))


def test_descriptor_channel_between_four_processes(lib):
    """SCM_RIGHTS plumbing used to hand VMM allocation handles to partner ranks: 4 processes, every pair swaps a pipe"""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    tag = int.from_bytes(os.urandom(7), "little")
    procs = [ctx.Process(target=_fdpass_worker, args=(r, 4, tag, q)) for r in range(4)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=60) for _ in procs)
    for p in procs:
        p.join(timeout=30)
    assert results == {0: 0, 1: 0, 2: 0, 3: 0}


def plan_tiles(lib, xmasks, zmasks, n_bits):
    n = len(xmasks)
    xs = (ctypes.c_uint64 * n)(*xmasks)
    zs = (ctypes.c_uint64 * n)(*zmasks)
    launch = (ctypes.c_int32 * n)()
    masks = (ctypes.c_uint64 * 256)()
    n_launch = ctypes.c_size_t()
    st = lib.pqb_host_plan_pauli_tiles(xs, zs, ctypes.c_size_t(n), ctypes.c_int(n_bits), launch, masks, ctypes.c_size_t(256),
                                       ctypes.byref(n_launch))
    assert st == 0
    return list(launch), [masks[i] for i in range(n_launch.value)]


def test_pauli_tile_planner_covers_every_term_once(lib):
    """every term is applied by exactly one launch whose tile bits contain its whole X/Y support (or is declared wide), a
    tile never has more than 11 bits, and the 28-qubit TFIM needs 3 tile-bit sets"""
    n = 28
    xm = [0] * (n - 1) + [1 << i for i in range(n)]
    zm = [(1 << i) | (1 << (i + 1)) for i in range(n - 1)] + [0] * n
    launch, masks = plan_tiles(lib, xm, zm, n)
    assert len(masks) == 3 and all(bin(m).count("1") == 11 for m in masks)
    for x, l in zip(xm, launch):
        assert 0 <= l < len(masks) and (x & ~masks[l]) == 0
    # random strings, some wider than a tile, more than 64 per set (several launches share one set of tile bits)
    rng = np.random.default_rng(9)
    n = 20
    xm, zm = [], []
    for _ in range(300):
        w = int(rng.integers(0, 6))
        xm.append(int(sum(1 << int(q) for q in rng.permutation(n)[:w])))
        zm.append(int(rng.integers(0, 1 << n)))
    xm += [(1 << 13) - 1, ((1 << 14) - 1) << 3]  # 13 and 14 flipped qubits: no 11-bit tile holds them
    zm += [0, 5]
    launch, masks = plan_tiles(lib, xm, zm, n)
    assert launch[-1] == -1 and launch[-2] == -1
    for x, l in zip(xm[:-2], launch[:-2]):
        assert 0 <= l < len(masks) and (x & ~masks[l]) == 0 and bin(masks[l]).count("1") <= 11
    # a state smaller than a tile: one launch, every bit a tile bit
    launch, masks = plan_tiles(lib, [1, 6, 0], [0, 1, 7], 3)
    assert masks == [7] and launch == [0, 0, 0]


_BLOCK_PLANNER_SCRIPT = r"""
import ctypes, os, sys
sys.path.insert(0, %r)
from tests.test_host_logic import plan_tiles
from projectq_b200 import _build
lib = ctypes.CDLL(_build.LIB)
n, B = 28, 20
xm = [0] * (n - 1) + [1 << i for i in range(n)]
zm = [(1 << i) | (1 << (i + 1)) for i in range(n - 1)] + [0] * n
launch, masks = plan_tiles(lib, xm, zm, n)
assert len(masks) == 3 and all(bin(m).count("1") == 11 for m in masks), masks
inside = [m for m in masks if m >> B == 0]
assert len(inside) == 2 and masks[:2] == inside, [hex(m) for m in masks]      # the sets below the block bits come first
for x, l in zip(xm, launch):
    assert 0 <= l < len(masks) and (x & ~masks[l]) == 0
    if x and x >> B == 0:
        assert masks[l] >> B == 0                                             # ... and take every term that fits below them
print("block planner OK")
"""


def test_pauli_tile_planner_orders_sets_for_block_fusion():
    """PQB_PAULI_BLOCK_BITS (opt-in fused execution): the tile-bit sets that lie below the block bits are planned first and
    take every term whose X/Y support fits below them; the rest of the cover is unchanged (3 sets for the 28-qubit TFIM)"""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", _BLOCK_PLANNER_SCRIPT % root], capture_output=True, text=True, timeout=120,
                         env=dict(os.environ, PQB_PAULI_BLOCK_BITS="20"), cwd=root)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "block planner OK" in out.stdout
