"""CPU restatement (NumPy) of ProjectQ's C++ state-vector simulator — TEST INFRASTRUCTURE ONLY.

This module is the *oracle* for the B200 engine: a readable restatement of the reference algorithm
(`projectq/backends/_sim/_cppkernels/simulator.hpp`, `fusion.hpp`, `nointrin/kernel*.hpp`) used by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg as the checker.  Nothing
under ``projectq_b200/`` may import it; the product path fails loudly when its CUDA library is missing.

Parity pinning: ``tests/test_oracle_pinned.py`` checks this restatement against
  * the compiled, unmodified reference (`oracle/_ref/_cppsim*.so`, built by `oracle/Makefile`) on seeded
    random circuits, and
  * the golden vectors generated from that reference (`tests/golden/*.json`, generator
    `tests/golden/make_golden.py`) and the known answers held by the reference's own tests
    (`_simulator_test.py:744` etc., restated in `tests/test_gpu_parity.py::test_reference_math_kats` and
    `tests/test_oracle_pinned.py`).

Each method cites the reference lines it follows (paths relative to /root/reference/projectq/backends/_sim).
Gate fusion is *not* restated: any fusion policy is amplitude-equivalent (simulator.hpp:204-222 only
batches the same products), so gates are applied one by one here.
"""

from __future__ import annotations

import cmath
import math

import numpy as np


class MT19937:
    """std::mt19937 seeded with one 32-bit value (libstdc++ ``mersenne_twister_engine::seed(value)``)."""

    def __init__(self, seed: int):
        mt = [0] * 624
        mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.mt = mt
        self.idx = 624

    def _twist(self):
        mt = self.mt
        for i in range(624):
            y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
            v = mt[(i + 397) % 624] ^ (y >> 1)
            if y & 1:
                v ^= 0x9908B0DF
            mt[i] = v
        self.idx = 0

    def next_u32(self) -> int:
        if self.idx >= 624:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def uniform01(self) -> float:
        """``std::uniform_real_distribution<double>(0,1)`` on mt19937 = generate_canonical<double,53>:
        two 32-bit draws, (u0 + u1*2^32) / 2^64, clamped below 1 (bits/random.tcc, libstdc++ 13)."""
        u0 = self.next_u32()
        u1 = self.next_u32()
        r = (float(u0) + float(u1) * 4294967296.0) / 18446744073709551616.0
        if r >= 1.0:
            r = math.nextafter(1.0, 0.0)
        return r


class OracleSimulator:
    """Same method surface as the reference's ``_cppsim.Simulator`` (_cppsim.cpp:43-67)."""

    def __init__(self, seed: int = 1):
        # simulator.hpp:48-53 — one amplitude, |> = 1, seeded mt19937
        self.vec = np.ones(1, dtype=np.complex128)
        self.map: dict[int, int] = {}
        self.rng = MT19937(seed)

    # ---- helpers -------------------------------------------------------------------------------
    @property
    def n(self) -> int:
        return len(self.map)

    def _mask(self, ids) -> int:
        m = 0
        for q in ids:
            m |= 1 << self.map[q]
        return m

    def _idx(self) -> np.ndarray:
        return np.arange(self.vec.size, dtype=np.uint64)

    # ---- allocation ----------------------------------------------------------------------------
    def allocate_qubit(self, qid: int):
        # simulator.hpp:55-74 — new qubit becomes the new most-significant bit; upper half zero
        if qid in self.map:
            raise RuntimeError("AllocateQubit: ID already exists. Qubit IDs should be unique.")
        self.map[qid] = self.n
        new = np.zeros(self.vec.size * 2, dtype=np.complex128)
        new[: self.vec.size] = self.vec
        self.vec = new

    def _bit_presence(self, qid: int, tol: float):
        pos = self.map[qid]
        nrm = self.vec.real**2 + self.vec.imag**2
        bit = (self._idx() >> np.uint64(pos)) & np.uint64(1)
        down = bool(np.any(nrm[bit == 0] > tol))  # "any amplitude with bit=0 present"
        up = bool(np.any(nrm[bit == 1] > tol))
        return down, up

    def is_classical(self, qid: int, tol: float = 1e-12) -> bool:
        # simulator.hpp:93-108
        down, up = self._bit_presence(qid, tol)
        return down != up

    def get_classical_value(self, qid: int, tol: float = 1e-12) -> bool:
        # simulator.hpp:76-91 — first amplitude above tol (scan order i, then i+delta) decides
        pos = self.map[qid]
        delta = 1 << pos
        nrm = self.vec.real**2 + self.vec.imag**2
        big = np.nonzero(nrm > tol)[0]
        if big.size == 0:
            raise AssertionError("no amplitude above tolerance")
        # reference scan order key: (block, j, half) where i = block*2*delta + half*delta + j
        blk = big // (2 * delta)
        j = big % delta
        half = (big // delta) & 1
        order = np.lexsort((half, j, blk))
        return bool(half[order[0]])

    def deallocate_qubit(self, qid: int):
        # simulator.hpp:194-202 + collapse_vector(shrink=true) :110-143
        if not self.is_classical(qid):
            raise RuntimeError(
                "Error: Qubit has not been measured / uncomputed! There is most likely a bug in your code."
            )
        value = self.get_classical_value(qid)
        pos = self.map[qid]
        bit = (self._idx() >> np.uint64(pos)) & np.uint64(1)
        self.vec = self.vec[bit == (1 if value else 0)].copy()  # index order kept, no renormalisation
        for k in list(self.map):
            if self.map[k] > pos:
                self.map[k] -= 1
        del self.map[qid]

    # ---- gates ---------------------------------------------------------------------------------
    def apply_controlled_gate(self, m, ids, ctrl):
        # caller-level contract of simulator.hpp:204-222 + run() :487-527 + nointrin/kernel2.hpp:16-28:
        # matrix bit l <-> ids[l]; psi'[I + sum bit_l(r) d_l] = sum_c M[r][c] psi[I + sum bit_l(c) d_l]
        # for all base I with target bits clear and all control bits set.
        m = np.asarray(m, dtype=np.complex128)
        k = len(ids)
        assert m.shape == (1 << k, 1 << k)
        pos = [self.map[q] for q in ids]
        cmask = self._mask(ctrl)
        tmask = sum(1 << p for p in pos)
        idx = self._idx()
        base = idx[((idx & np.uint64(tmask)) == 0) & ((idx & np.uint64(cmask)) == np.uint64(cmask))]
        offs = np.zeros(1 << k, dtype=np.uint64)
        for j in range(1 << k):
            o = 0
            for l in range(k):
                if (j >> l) & 1:
                    o |= 1 << pos[l]
            offs[j] = o
        gather = base[:, None] + offs[None, :]  # (groups, 2^k)
        v = self.vec[gather]
        self.vec[gather] = v @ m.T

    def run(self):
        # simulator.hpp:487-527 — nothing is queued in the oracle
        return None

    # ---- measurement ---------------------------------------------------------------------------
    def measure_qubits(self, ids):
        # simulator.hpp:145-186
        rnd = self.rng.uniform01()
        # sequential running sum, exactly like the reference's while loop (:156-158)
        nrm = self.vec.real**2 + self.vec.imag**2
        csum = np.cumsum(nrm)  # numpy cumsum is a sequential left-to-right double sum
        pick = int(np.searchsorted(csum, rnd, side="left"))  # first index with csum >= rnd
        if pick >= self.vec.size:
            pick = self.vec.size - 1
        res = []
        mask = 0
        val = 0
        for q in ids:
            p = self.map[q]
            b = (pick >> p) & 1
            res.append(bool(b))
            mask |= 1 << p
            val |= b << p
        keep = (self._idx() & np.uint64(mask)) == np.uint64(val)
        self.vec[~keep] = 0.0
        norm = float(np.sum(nrm[keep]))
        self.vec *= 1.0 / math.sqrt(norm)
        return res

    def collapse_wavefunction(self, ids, values):
        # simulator.hpp:456-485
        if len(ids) != len(values):
            raise ValueError("collapse_wavefunction(): ids and values size mismatch")
        if any(q not in self.map for q in ids):
            raise RuntimeError(
                "collapse_wavefunction(): Unknown qubit id(s) provided. Try calling eng.flush() before "
                "invoking this function."
            )
        mask = 0
        val = 0
        for q, v in zip(ids, values):
            mask |= 1 << self.map[q]
            val |= (1 if v else 0) << self.map[q]
        keep = (self._idx() & np.uint64(mask)) == np.uint64(val)
        nrm = self.vec.real**2 + self.vec.imag**2
        prob = float(np.sum(nrm[keep]))
        if prob < 1e-12:
            raise RuntimeError("collapse_wavefunction(): Invalid collapse! Probability is ~0.")
        self.vec[~keep] = 0.0
        self.vec[keep] *= 1.0 / math.sqrt(prob)

    # ---- queries -------------------------------------------------------------------------------
    def get_probability(self, bits, ids) -> float:
        # simulator.hpp:352-368
        if any(q not in self.map for q in ids):
            raise RuntimeError(
                "get_probability(): Unknown qubit id. Please make sure you have called eng.flush()."
            )
        mask = 0
        val = 0
        for q, b in zip(ids, bits):
            mask |= 1 << self.map[q]
            val |= (1 if b else 0) << self.map[q]
        keep = (self._idx() & np.uint64(mask)) == np.uint64(val)
        nrm = self.vec.real**2 + self.vec.imag**2
        return float(np.sum(nrm[keep]))

    def get_amplitude(self, bits, ids) -> complex:
        # simulator.hpp:370-384 — ids must be a permutation of all allocated qubits
        chk = 0
        index = 0
        for q, b in zip(ids, bits):
            if q not in self.map:
                break
            chk |= 1 << self.map[q]
            index |= (1 if b else 0) << self.map[q]
        if chk + 1 != self.vec.size:
            raise RuntimeError(
                "The second argument to get_amplitude() must be a permutation of all allocated qubits. "
                "Please make sure you have called eng.flush()."
            )
        return complex(self.vec[index])

    def set_wavefunction(self, wf, ordering):
        # simulator.hpp:440-454
        if len(self.map) != len(ordering) or any(q not in self.map for q in ordering):
            raise RuntimeError(
                "set_wavefunction(): Invalid mapping provided. Please make sure all qubits have been "
                "allocated previously (call eng.flush())."
            )
        for i, q in enumerate(ordering):
            self.map[q] = i
        self.vec = np.array(wf, dtype=np.complex128).copy()

    def cheat(self):
        # simulator.hpp:529-532
        return dict(self.map), self.vec.copy()

    # ---- Pauli strings -------------------------------------------------------------------------
    def _pauli_apply(self, term, ids, vec):
        """(P vec)[j] = i^{nY} (-1)^{popcount(s & zmask)} vec[s], s = j ^ xmask  (apply_term,
        simulator.hpp:538-550 with the X/Y/Z matrices of :541-543)."""
        if self._has_repeat(term):
            return self._pauli_apply_seq(term, ids, vec)
        xmask = 0
        zmask = 0
        ny = 0
        for local, op in term:
            p = self.map[ids[local]]
            if op == "X":
                xmask |= 1 << p
            elif op == "Z":
                zmask |= 1 << p
            elif op == "Y":
                xmask |= 1 << p
                zmask |= 1 << p
                ny += 1
            else:
                raise ValueError(op)
        return self._pauli_masks(xmask, zmask, ny, vec)

    @staticmethod
    def _has_repeat(term):
        seen = set()
        for local, _ in term:
            if local in seen:
                return True
            seen.add(local)
        return False

    def _pauli_masks(self, xmask, zmask, ny, vec):
        idx = self._idx()
        src = idx ^ np.uint64(xmask)
        par = np.zeros(vec.size, dtype=np.uint64)
        z = src & np.uint64(zmask)
        for b in range(max(1, self.n)):
            par ^= (z >> np.uint64(b)) & np.uint64(1)
        sign = 1.0 - 2.0 * par.astype(np.float64)
        return (1j**ny) * sign * vec[src]

    def _pauli_apply_seq(self, term, ids, vec):
        # general fall-back: successive one-qubit gates exactly as apply_term queues them
        out = vec.copy()
        mats = {
            "X": np.array([[0, 1], [1, 0]], dtype=np.complex128),
            "Y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
            "Z": np.array([[1, 0], [0, -1]], dtype=np.complex128),
        }
        saved = self.vec
        self.vec = out
        for local, op in term:
            self.apply_controlled_gate(mats[op], [ids[local]], [])
        out = self.vec
        self.vec = saved
        return out

    def get_expectation_value(self, terms, ids) -> float:
        # simulator.hpp:292-322 — sum_t c_t * Re<psi|P_t|psi>, c_t real
        e = 0.0
        for term, coeff in terms:
            if isinstance(coeff, complex):
                raise TypeError("get_expectation_value(): coefficients must be real")
            pv = self._pauli_apply(term, ids, self.vec)
            delta = float(np.sum(self.vec.real * pv.real + self.vec.imag * pv.imag))
            e += coeff * delta
        return e

    def apply_qubit_operator(self, terms, ids):
        # simulator.hpp:324-350 — psi <- sum_t c_t P_t psi, no renormalisation
        new = np.zeros_like(self.vec)
        for term, coeff in terms:
            new += complex(coeff) * self._pauli_apply(term, ids, self.vec)
        self.vec = new

    def emulate_time_evolution(self, terms, time, ids, ctrl):
        # simulator.hpp:386-438
        tr = 0.0
        op_nrm = 0.0
        td = []
        for term, coeff in terms:
            if isinstance(coeff, complex):
                raise TypeError("emulate_time_evolution(): coefficients must be real")
            if len(term) == 0:
                tr += coeff
            else:
                td.append((term, coeff))
                op_nrm += abs(coeff)
        s = int(abs(time) * op_nrm + 1.0)
        correction = cmath.exp(-time * 1j * tr / float(s))
        cmask = self._mask(ctrl)
        sel = (self._idx() & np.uint64(cmask)) == np.uint64(cmask)
        out = self.vec.copy()
        cur = self.vec.copy()
        for _ in range(s):
            nrm_change = 1.0
            k = 0
            v = cur
            while nrm_change > 1e-12:
                coeff = (-time * 1j) / float(s * (k + 1))
                upd = np.zeros_like(v)
                for term, c in td:
                    upd += self._pauli_apply(term, ids, v) * c
                upd *= coeff
                v = upd
                out[sel] += upd[sel]
                nrm_change = math.sqrt(float(np.sum(upd[sel].real ** 2 + upd[sel].imag ** 2)))
                k += 1
            out[sel] *= correction
            cur = out.copy()
        self.vec = cur

    # ---- emulated arithmetic -------------------------------------------------------------------
    def emulate_math(self, f, quregs, ctrl):
        # simulator.hpp:224-269 — new[pi(i)] += psi[i]; identity where the control bits are not all set
        cmask = self._mask(ctrl)
        pos = [[self.map[q] for q in reg] for reg in quregs]
        new = np.zeros_like(self.vec)
        cache: dict[tuple, list] = {}
        for i in range(self.vec.size):
            if (i & cmask) == cmask:
                res = tuple(sum(((i >> p) & 1) << b for b, p in enumerate(reg)) for reg in pos)
                if res not in cache:
                    cache[res] = list(f(list(res)))
                out = cache[res]
                ni = i
                for reg, r in zip(pos, out):
                    for b, p in enumerate(reg):
                        if ((ni >> p) & 1) != ((int(r) >> b) & 1):  # Python >> on negatives = two's complement
                            ni ^= 1 << p
                new[ni] += self.vec[i]
            else:
                new[i] += self.vec[i]
        self.vec = new

    @staticmethod
    def _c_int(x: int) -> int:
        """wrap to a C ``int`` (32-bit two's complement), the type used at simulator.hpp:244."""
        x &= 0xFFFFFFFF
        return x - (1 << 32) if x & 0x80000000 else x

    @staticmethod
    def _c_mod(a: int, n: int) -> int:
        """C ``%``: truncation toward zero."""
        return int(math.fmod(a, n)) if abs(a) < 2**52 else a - n * int(a / n)

    def emulate_math_addConstant(self, a, quregs, ctrl):
        # simulator.hpp:271-276
        self.emulate_math(lambda r: [self._c_int(x + a) for x in r], quregs, ctrl)

    def emulate_math_addConstantModN(self, a, N, quregs, ctrl):
        # simulator.hpp:278-283
        self.emulate_math(lambda r: [self._c_mod(self._c_int(x + a), N) for x in r], quregs, ctrl)

    def emulate_math_multiplyByConstantModN(self, a, N, quregs, ctrl):
        # simulator.hpp:285-290 (32-bit int product; parity defined for x*a < 2^31)
        self.emulate_math(lambda r: [self._c_mod(self._c_int(x * a), N) for x in r], quregs, ctrl)
