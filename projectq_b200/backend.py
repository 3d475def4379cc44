"""Loader of the native seam (pybind11 shim over the C ABI).  Fails loudly when the CUDA extension is missing."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpqb200.so")

try:
    from . import _pqb_shim
except ImportError as exc:  # no silent fallback: the reference would drop to _pysim here (_simulator.py:37-43), we do not
    raise ImportError(
        "projectq_b200: the CUDA extension is not built (run `python -m projectq_b200._build` or "
        "`__graft_entry__.build()`); there is no CPU fallback. Cause: %s" % (exc,)
    ) from exc

SimulatorBackend = _pqb_shim.Simulator
nccl_unique_id = _pqb_shim.nccl_unique_id
version = _pqb_shim.version
selftest_exchange = _pqb_shim.selftest_exchange


def load_c_abi():
    """ctypes handle on libpqb200.so (used by the symbol-export test and by bindings that skip pybind11)."""
    return ctypes.CDLL(LIB_PATH)


class DeviceStateView:
    """Zero-copy handle on a backend's state in HBM: exposes ``__cuda_array_interface__`` (complex128, one entry per local
    amplitude) so that ``torch.as_tensor(view, device="cuda")``, ``cupy.asarray(view)`` or numba wrap the memory without a
    copy.  ``layout[p]`` is the physical bit of logical bit position p (local bit < 64, else 64 + rank bit).  Keeps the
    backend alive; the memory is only valid until the next call that allocates, deallocates or remaps qubits."""

    def __init__(self, backend):
        info = backend.state_view()
        self._backend = backend
        self.__cuda_array_interface__ = info["cuda_array_interface"]
        self.layout = list(info["layout"])
