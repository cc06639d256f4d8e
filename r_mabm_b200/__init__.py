"""r_mabm_b200 -- B200-native batched CATS economy engine behind the rmabm `Cats` API.

Hot path only: the per-period model step that `rmabm.environments.Cats.step` delegates to
ABCredit.jl (/root/reference/rmabm/environments/cats.py:355-410), as hand-written CUDA for sm_100a
behind a C-ABI (include/cats_b200.h).  See docs/DESIGN.md.
"""
from .params import DEFAULT_PARAMETERS, PARAM_KEYS, load_parameters  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):  # lazy: importing the package must not need torch / CUDA / the .so
    if name == "BatchedCats":
        from .engine import BatchedCats
        return BatchedCats
    if name in ("Cats", "CatsLog"):
        from .environments import cats as _c
        return getattr(_c, name)
    if name in ("Simulation",):
        from .simulation import Simulation
        return Simulation
    if name in ("Logger",):
        from .logger import Logger
        return Logger
    raise AttributeError(name)
