"""B200-native state-vector engine behind ProjectQ's ``Simulator`` backend.

* ``projectq_b200.backend.SimulatorBackend`` — the native seam: an object with the method surface of the reference's
  ``_cppsim.Simulator`` (reference: projectq/backends/_sim/_cppsim.cpp:43-67), implemented by hand-written sm_100a CUDA
  kernels behind the C ABI of ``include/pqb200.h``.
* ``projectq_b200.Simulator`` — the ProjectQ compiler engine (reference: projectq/backends/_sim/_simulator.py:46-438);
  importable only where the ``projectq`` package is installed, because it subclasses ``projectq.cengines.BasicEngine``.
* ``projectq_b200.UnitarySimulator`` — the unitary-recording backend (reference: projectq/backends/_unitary.py:64-285) on
  the same kernels: the 2^n columns of the unitary are 2^n state vectors side by side.

There is no CPU or ``_pysim`` fallback: importing the backend without the built CUDA library raises ImportError and
constructing it without a CUDA device raises RuntimeError.
"""

__all__ = ["Simulator", "SimulatorBackend", "UnitarySimulator"]


def __getattr__(name):
    if name == "SimulatorBackend":
        from .backend import SimulatorBackend

        return SimulatorBackend
    if name == "Simulator":
        from ._simulator import Simulator

        return Simulator
    if name == "UnitarySimulator":
        from ._unitary import UnitarySimulator

        return UnitarySimulator
    raise AttributeError(name)
