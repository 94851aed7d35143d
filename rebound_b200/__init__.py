"""rebound_b200 -- B200-native force / collision / kick-drift hot path behind REBOUND's C API.

`abi` and `ics` are pure Python; everything that computes goes through the CUDA library
(rebound_b200/librebound_b200.so) and raises if that library is missing -- there is no CPU path.
"""
from . import abi, ics  # noqa: F401

__all__ = ["abi", "ics", "Simulation", "load_library"]


def __getattr__(name):
    if name in ("Simulation", "load_library", "LibraryMissing", "ReboundCudaError"):
        from . import simulation

        return getattr(simulation, name)
    raise AttributeError(name)
