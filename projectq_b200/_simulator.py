"""ProjectQ compiler engine that simulates on the B200 state-vector engine.

Drop-in for ``projectq.backends.Simulator`` (reference: projectq/backends/_sim/_simulator.py:46-438): same constructor
arguments, same user API (``get_expectation_value``, ``apply_qubit_operator``, ``get_probability``, ``get_amplitude``,
``set_wavefunction``, ``collapse_wavefunction``, ``cheat``), same command handling (Measure / Allocate / Deallocate /
BasicMathGate / TimeEvolution / matrix gates up to 5 qubits with positive controls), same exceptions.  The native object
behind ``self._simulator`` is the CUDA backend; there is no ``_pysim`` fallback.

This module needs the ``projectq`` package (it subclasses ``projectq.cengines.BasicEngine``, which ``MainEngine`` insists
on, reference: cengines/_main.py:136-145).  The native seam itself (``projectq_b200.backend``) does not.
"""
import math
import random

from projectq.cengines import BasicEngine
from projectq.meta import LogicalQubitIDTag, get_control_count, has_negative_control
from projectq.ops import Allocate, BasicMathGate, Deallocate, FlushGate, Measure, TimeEvolution
from projectq.types import WeakQubitRef

from .backend import SimulatorBackend

_MAX_GATE_QUBITS = 5


def _ids(qubits):
    return [qb.id for qb in qubits]


class Simulator(BasicEngine):
    """Simulates a quantum computer on one (or several) B200 GPUs through hand-written CUDA kernels."""

    def __init__(self, gate_fusion=False, rnd_seed=None, **backend_options):
        """
        Args:
            gate_fusion (bool): buffer gates and apply them as fused dense passes of up to 5 qubits when the queue is
                flushed (reference: _simulator.py:58-84; there it "may or may not be beneficial", here it always is).
            rnd_seed (int): seed of the measurement RNG; ``random.randint(0, 4294967295)`` by default as in the reference.
            backend_options: extra keyword arguments for the native backend (``device``, ``fusion_max_qubits``,
                ``rank``, ``world_size``, ``nccl_unique_id``, ``reserve_qubits``); the positional surface is unchanged.
        """
        if rnd_seed is None:
            rnd_seed = random.randint(0, 4294967295)
        super().__init__()
        self._simulator = SimulatorBackend(rnd_seed, **backend_options)
        self._gate_fusion = gate_fusion

    # ------------------------------------------------------------------------------------------------------------
    def is_available(self, cmd):
        """Matrix gates on up to 5 target qubits with arbitrary positive controls, Measure/Allocate/Deallocate, math gates
        and TimeEvolution are simulated directly (reference: _simulator.py:86-116)."""
        if has_negative_control(cmd):
            return False
        gate = cmd.gate
        if gate == Measure or gate == Allocate or gate == Deallocate:
            return True
        if isinstance(gate, (BasicMathGate, TimeEvolution)):
            return True
        try:
            return len(gate.matrix) <= 2**_MAX_GATE_QUBITS
        except AttributeError:
            return False

    def _convert_logical_to_mapped_qureg(self, qureg):
        """Translate logical qubits to mapped ids when the compiler chain has a mapper (reference: _simulator.py:118-134)."""
        mapper = self.main_engine.mapper
        if mapper is None:
            return qureg
        mapped = []
        for qubit in qureg:
            try:
                new_id = mapper.current_mapping[qubit.id]
            except KeyError:
                raise RuntimeError("Unknown qubit id. Please make sure you have called eng.flush().") from None
            mapped.append(WeakQubitRef(qubit.engine, new_id))
        return mapped

    @staticmethod
    def _operator_terms(qubit_operator, num_qubits):
        terms = []
        for term, coefficient in qubit_operator.terms.items():
            if term != () and term[-1][0] >= num_qubits:
                raise Exception("qubit_operator acts on more qubits than contained in the qureg.")
            terms.append((list(term), coefficient))
        return terms

    # ---- user API --------------------------------------------------------------------------------------------
    def get_expectation_value(self, qubit_operator, qureg):
        """<psi| qubit_operator |psi> for the register ``qureg`` (reference: _simulator.py:136-167)."""
        qureg = self._convert_logical_to_mapped_qureg(qureg)
        terms = self._operator_terms(qubit_operator, len(qureg))
        return self._simulator.get_expectation_value(terms, _ids(qureg))

    def apply_qubit_operator(self, qubit_operator, qureg):
        """psi <- qubit_operator psi, not renormalised (reference: _simulator.py:169-199)."""
        qureg = self._convert_logical_to_mapped_qureg(qureg)
        terms = self._operator_terms(qubit_operator, len(qureg))
        return self._simulator.apply_qubit_operator(terms, _ids(qureg))

    def get_probability(self, bit_string, qureg):
        """Probability of measuring ``bit_string`` (list of bool/int or a '0101' string) on ``qureg``
        (reference: _simulator.py:201-222)."""
        qureg = self._convert_logical_to_mapped_qureg(qureg)
        return self._simulator.get_probability([bool(int(b)) for b in bit_string], _ids(qureg))

    def get_amplitude(self, bit_string, qureg):
        """Amplitude of the basis state ``bit_string``; ``qureg`` must hold all allocated qubits
        (reference: _simulator.py:224-248)."""
        qureg = self._convert_logical_to_mapped_qureg(qureg)
        return self._simulator.get_amplitude([bool(int(b)) for b in bit_string], _ids(qureg))

    def set_wavefunction(self, wavefunction, qureg):
        """Overwrite the state; the simulator adopts the qubit ordering of ``qureg`` (reference: _simulator.py:250-274)."""
        qureg = self._convert_logical_to_mapped_qureg(qureg)
        self._simulator.set_wavefunction(wavefunction, _ids(qureg))

    def collapse_wavefunction(self, qureg, values):
        """Project ``qureg`` onto ``values`` and renormalise; RuntimeError if the outcome has probability ~0
        (reference: _simulator.py:276-299)."""
        qureg = self._convert_logical_to_mapped_qureg(qureg)
        return self._simulator.collapse_wavefunction(_ids(qureg), [bool(int(v)) for v in values])

    def cheat(self):
        """(id -> bit position dict, state vector as a NumPy complex128 array) (reference: _simulator.py:301-322)."""
        return self._simulator.cheat()

    # ---- additions: the state as data -------------------------------------------------------
    def save_state(self, path_prefix):
        """Write the state, the qubit map and the RNG position to ``<path_prefix>.rank<r>of<w>.pqbs`` (one file per GPU of a
        sharded run; nothing is gathered).  Call ``eng.flush()`` first."""
        self._simulator.save_state(str(path_prefix))

    def load_state(self, path_prefix):
        """Replace the simulator's qubits, state and RNG position by a checkpoint written with ``save_state``.  The program
        must hold qubits with the same ids (allocate the same registers in the same order, ``eng.flush()``, then load)."""
        self._simulator.load_state(str(path_prefix))

    # ---- command handling --------------------------------------------------------------------------------------
    def _handle(self, cmd):
        """Dispatch one command to the native backend (reference: _simulator.py:324-420)."""
        gate = cmd.gate
        if gate == Measure:
            if get_control_count(cmd) != 0:
                raise ValueError('Cannot have control qubits with a measurement gate!')
            qubits = [qb for qureg in cmd.qubits for qb in qureg]
            outcome = self._simulator.measure_qubits(_ids(qubits))
            logical_tag = None
            for tag in cmd.tags:
                if isinstance(tag, LogicalQubitIDTag):
                    logical_tag = tag
            for qb, bit in zip(qubits, outcome):
                if logical_tag is not None:  # a mapper relabelled the qubit: report under its logical id
                    qb = WeakQubitRef(qb.engine, logical_tag.logical_qubit_id)
                self.main_engine.set_measurement_result(qb, bit)
        elif gate == Allocate:
            self._simulator.allocate_qubit(cmd.qubits[0][0].id)
        elif gate == Deallocate:
            self._simulator.deallocate_qubit(cmd.qubits[0][0].id)
        elif isinstance(gate, BasicMathGate):
            self._handle_math(cmd)
        elif isinstance(gate, TimeEvolution):
            terms = [(list(term), coefficient) for term, coefficient in gate.hamiltonian.terms.items()]
            self._simulator.emulate_time_evolution(terms, gate.time, _ids(cmd.qubits[0]), _ids(cmd.control_qubits))
        else:
            matrix, ids, ctrl = self._matrix_gate(cmd)
            native = self._simulator
            if hasattr(native, "apply_gate_list"):
                native.apply_gate_list([(matrix, ids, ctrl)], self._gate_fusion)
            else:  # a reference-style native object swapped in (the reference's tests do that, _simulator_test.py:80-93)
                native.apply_controlled_gate(matrix.tolist() if hasattr(matrix, "tolist") else matrix, ids, ctrl)
                if not self._gate_fusion:
                    native.run()

    @staticmethod
    def _matrix_gate(cmd):
        """(matrix, target ids, control ids) of a matrix-gate command, with the reference's checks and messages
        (reference: _simulator.py:399-420)"""
        gate = cmd.gate
        matrix = gate.matrix
        if len(matrix) > 2**_MAX_GATE_QUBITS:
            raise Exception(
                "This simulator only supports controlled k-qubit gates with k < 6!\nPlease add an auto-replacer"
                " engine to your list of compiler engines."
            )
        ids = [qb.id for qureg in cmd.qubits for qb in qureg]
        if 2 ** len(ids) != len(matrix):
            raise Exception(
                f"Simulator: Error applying {str(gate)} gate: {int(math.log(len(matrix), 2))}-qubit"
                f" gate applied to {len(ids)} qubits."
            )
        return matrix, ids, _ids(cmd.control_qubits)

    def _handle_math(self, cmd):
        """Emulated arithmetic: closed-form kernels for the three constant-math gates, a lookup table built from the
        gate's Python function for everything else (reference: _simulator.py:361-397)."""
        from projectq.libs.math import AddConstant, AddConstantModN, MultiplyByConstantModN

        gate = cmd.gate
        quregs = [_ids(qureg) for qureg in cmd.qubits]
        ctrl = _ids(cmd.control_qubits)
        if isinstance(gate, AddConstant):
            self._simulator.emulate_math_addConstant(gate.a, quregs, ctrl)
        elif isinstance(gate, AddConstantModN):
            self._simulator.emulate_math_addConstantModN(gate.a, gate.N, quregs, ctrl)
        elif isinstance(gate, MultiplyByConstantModN):
            self._simulator.emulate_math_multiplyByConstantModN(gate.a, gate.N, quregs, ctrl)
        else:
            self._simulator.emulate_math(gate.get_math_function(cmd.qubits), quregs, ctrl)

    def receive(self, command_list):
        """Simulate the commands, then pass them on if this is not the last engine (reference: _simulator.py:422-438).

        The reference makes one native call (and one ``matrix.tolist()``) per command.  Here consecutive matrix gates of a
        command list are handed to the native backend in ONE call (``apply_gate_list`` -> ``pqb_apply_gate_stream``), in
        program order, before any other kind of command is handled and before this method returns, so every observable
        (Measure, user API calls, an exception raised by a later command) sees the same simulator state as in the
        reference."""
        native = self._simulator
        batched = hasattr(native, "apply_gate_list")
        batch = []

        def flush_batch():
            if batch:
                native.apply_gate_list(batch, self._gate_fusion)
                del batch[:]

        for cmd in command_list:
            gate = cmd.gate
            if gate == FlushGate():
                flush_batch()
                native.run()
            elif (batched and gate != Measure and gate != Allocate and gate != Deallocate
                  and not isinstance(gate, (BasicMathGate, TimeEvolution))):
                try:
                    batch.append(self._matrix_gate(cmd))
                except Exception:
                    flush_batch()  # everything before the offending command has been applied, as in the reference
                    raise
            else:
                flush_batch()
                self._handle(cmd)
            if not self.is_last_engine:
                self.send([cmd])
        flush_batch()
