"""ProjectQ backend that records the unitary of a circuit, computed on the B200 state-vector engine.

Drop-in for ``projectq.backends.UnitarySimulator`` (reference: projectq/backends/_unitary.py:64-285): same constructor,
``unitary`` / ``history`` properties, ``is_available``, command handling (Allocate / Deallocate / Measure / Flush / matrix
gates with positive controls), warnings and exceptions.

Where the reference multiplies a dense 2^n x 2^n matrix into the running unitary for every gate (O(8^n) per gate,
_unitary.py:236-245), the unitary here lives on the GPU as a "state" of 2n bits — n row bits and n column bits, i.e. the
2^n columns of U are 2^n state vectors stored side by side — and a gate on qubit q is the engine's ordinary fused dense
pass on row bit q, the column bits being spectators: one HBM sweep of 16 B x 4^n per fused pass.  The identity is prepared
exactly (an unnormalised |00> + |11> on every (row, column) pair: the dense kernels take any matrix, not only unitaries),
and allocating a qubit mid-circuit grows it the same way (U <- 1 (x) U, _unitary.py:198-199).

The small 2^n state vector that the reference keeps for measurements (``_state``, advanced by ``_flush`` and collapsed by
``measure_qubits`` with Python's ``random``) is kept on the host exactly as in the reference, quirks included: a second
flush after more gates applies the *accumulated* unitary to the already advanced state (_unitary.py:184-187).

One more quirk is reproduced because results must equal the reference's: for a multi-qubit gate the reference orders the
matrix index bits by the *positions* of the target qubits, not by the order in which the command lists them
(_qidmask, _unitary.py:31-61), unlike ``Simulator``.
"""
import math
import random
import warnings
from copy import deepcopy

import numpy as np

from projectq.cengines import BasicEngine
from projectq.meta import LogicalQubitIDTag, get_control_count, has_negative_control
from projectq.ops import AllocateQubitGate, DeallocateQubitGate, FlushGate, MeasureGate
from projectq.types import WeakQubitRef

from .backend import SimulatorBackend

_MAX_GATE_QUBITS = 5


class UnitarySimulator(BasicEngine):
    """Calculates the unitary transformation that represents the circuit processed so far, on the GPU."""

    def __init__(self, **backend_options):
        """
        Args:
            backend_options: extra keyword arguments for the native backend (``device``, ``fusion_max_qubits``); the
                reference's constructor takes no arguments.
        """
        super().__init__()
        self._backend_options = backend_options
        self._native = None     # SimulatorBackend holding the 4^n entries of the unitary
        self._rows = []         # native ids of the row bits, by qubit position
        self._next_native_id = 0
        self._qubit_map = {}    # qubit id -> position (reference: _unitary.py:92)
        self._num_qubits = 0
        self._is_valid = True
        self._is_flushed = False
        self._state = np.array([1], dtype=complex)
        self._history = []
        self._host_copy = None  # the unitary as an ndarray, valid until the next gate

    # ---- the unitary on the device --------------------------------------------------------------------------------
    def _add_pair(self):
        """U <- 1 (x) U: a new (row, column) bit pair in the unnormalised state |00> + |11>"""
        if self._native is None:
            self._native = SimulatorBackend(1, **self._backend_options)
        row, col = self._next_native_id, self._next_native_id + 1
        self._next_native_id += 2
        self._native.allocate_qubit(row)
        self._native.allocate_qubit(col)
        self._native.apply_controlled_gate(np.array([[1, 1], [1, -1]], dtype=complex), [col], [])
        self._native.apply_controlled_gate(np.array([[0, 1], [1, 0]], dtype=complex), [row], [col])
        self._rows.append(row)
        self._host_copy = None

    def _reset(self, n_qubits):
        """fresh identity on n_qubits (reference: _unitary.py:229-231)"""
        self._native = None
        self._rows = []
        self._next_native_id = 0
        for _ in range(n_qubits):
            self._add_pair()

    def _download(self):
        """the unitary as a 2^n x 2^n ndarray, U[r, c] with qubit position p <-> bit p of r and c"""
        if self._host_copy is not None:
            return self._host_copy
        n = len(self._rows)
        if n == 0:
            return [1]  # what the reference holds before the first allocation (_unitary.py:93)
        mapping, vec = self._native.cheat()
        mapping = dict(mapping)
        t = np.asarray(vec).reshape([2] * (2 * n))  # axis a <-> bit 2n-1-a of the flat index
        axis = lambda native_id: 2 * n - 1 - mapping[native_id]  # noqa: E731
        order = [axis(self._rows[p]) for p in reversed(range(n))] + [axis(self._rows[p] + 1) for p in reversed(range(n))]
        self._host_copy = np.ascontiguousarray(t.transpose(order)).reshape(1 << n, 1 << n)
        return self._host_copy

    # ---- reference surface ------------------------------------------------------------------------------------------
    @property
    def unitary(self):
        """A copy of the current unitary matrix (reference: _unitary.py:101-109)."""
        return deepcopy(self._download())

    @property
    def history(self):
        """Copies of all previous unitaries, separated by measurement / deallocation (reference: _unitary.py:111-123)."""
        return deepcopy(self._history)

    def is_available(self, cmd):
        """All gates with a matrix and positive controls, plus Allocate / Deallocate / Measure (reference:
        _unitary.py:125-151)."""
        if has_negative_control(cmd):
            return False
        if isinstance(cmd.gate, (AllocateQubitGate, DeallocateQubitGate, MeasureGate)):
            return True
        try:
            gate_mat = cmd.gate.matrix
            if len(gate_mat) > 2**6:
                warnings.warn(f"Potentially large matrix gate encountered! ({math.log2(len(gate_mat))} qubits)")
            return True
        except AttributeError:
            return False

    def receive(self, command_list):
        """Handle the commands, then send them on (reference: _unitary.py:153-172)."""
        for cmd in command_list:
            self._handle(cmd)
        if not self.is_last_engine:
            self.send(command_list)

    def _flush(self):
        """state <- U state, once per batch of gates (reference: _unitary.py:174-178)"""
        if not self._is_flushed:
            self._is_flushed = True
            self._state = np.asarray(self._download()) @ self._state

    def _handle(self, cmd):
        gate = cmd.gate
        if isinstance(gate, AllocateQubitGate):
            self._qubit_map[cmd.qubits[0][0].id] = self._num_qubits
            self._num_qubits += 1
            self._add_pair()
            self._state = np.concatenate([self._state, np.zeros(len(self._state), dtype=complex)])
        elif isinstance(gate, DeallocateQubitGate):
            pos = self._qubit_map[cmd.qubits[0][0].id]
            self._qubit_map = {key: value - 1 if value > pos else value for key, value in self._qubit_map.items()}
            self._num_qubits -= 1
            self._is_valid = False
        elif isinstance(gate, MeasureGate):
            self._is_valid = False
            if not self._is_flushed:
                raise RuntimeError(
                    'Please make sure all previous gates are flushed before measurement so the state gets updated'
                )
            if get_control_count(cmd) != 0:
                raise ValueError('Cannot have control qubits with a measurement gate!')
            all_qubits = [qb for qr in cmd.qubits for qb in qr]
            measurements = self.measure_qubits([qb.id for qb in all_qubits])
            for qb, res in zip(all_qubits, measurements):
                for tag in cmd.tags:  # a mapper assigned a different logical id
                    if isinstance(tag, LogicalQubitIDTag):
                        qb = WeakQubitRef(qb.engine, tag.logical_qubit_id)
                        break
                self.main_engine.set_measurement_result(qb, res)
        elif isinstance(gate, FlushGate):
            self._flush()
        else:
            if not self._is_valid:
                self._flush()
                warnings.warn(
                    "Processing of other gates after a qubit deallocation or measurement will reset the unitary,"
                    "previous unitary can be accessed in history"
                )
                self._history.append(self._download())
                self._reset(self._num_qubits)
                self._state = np.array([1] + ([0] * (2**self._num_qubits - 1)), dtype=complex)
                self._is_valid = True
            self._is_flushed = False
            self._host_copy = None
            matrix = gate.matrix
            positions = [self._qubit_map[qb.id] for qr in cmd.qubits for qb in qr]
            if 2 ** len(positions) != len(matrix):
                raise ValueError(f"UnitarySimulator: {len(matrix)}x{len(matrix)} matrix applied to {len(positions)} qubits")
            if len(positions) > _MAX_GATE_QUBITS:
                raise Exception(
                    "This backend supports gates on up to 5 target qubits; add an auto-replacer engine for wider gates."
                )
            # matrix index bit l <-> the target with the l-th lowest position (reference: _qidmask, _unitary.py:31-61)
            targets = [self._rows[p] for p in sorted(positions)]
            controls = [self._rows[self._qubit_map[qb.id]] for qb in cmd.control_qubits]
            self._native.apply_controlled_gate(np.asarray(matrix, dtype=complex), targets, controls)

    def measure_qubits(self, ids):
        """Measure on the host-side state with Python's RNG and collapse it (reference: _unitary.py:247-285)."""
        random_outcome = random.random()
        cdf = np.cumsum(np.abs(self._state) ** 2)
        # first index whose running sum reaches the draw, the last index if none does (the reference's while loop)
        hit = np.nonzero(cdf >= random_outcome)[0]
        i_picked = int(hit[0]) if len(hit) else len(self._state) - 1
        pos = [self._qubit_map[ID] for ID in ids]
        res = [((i_picked >> p) & 1) == 1 for p in pos]
        mask = 0
        val = 0
        for r, p in zip(res, pos):
            mask |= 1 << p
            val |= int(r) << p
        idx = np.arange(len(self._state))
        keep = (idx & mask) == val
        self._state = np.where(keep, self._state, 0.0)
        self._state = self._state * (1.0 / np.sqrt(np.sum(np.abs(self._state) ** 2)))
        return res
