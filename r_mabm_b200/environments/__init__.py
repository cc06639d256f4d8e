"""Environments: `Cats`, `CatsLog` (reference: rmabm/environments/__init__.py:1-13).

The gym ids "Cats" and "CatsLog" are registered only when gymnasium is importable; the classes work
without it (r_mabm_b200/environments/spaces.py provides the space stand-ins).
"""
from .cats import Cats, CatsLog

try:  # pragma: no cover - gymnasium is optional
    from gymnasium import register

    register(id="Cats", entry_point="r_mabm_b200.environments:Cats")
    register(id="CatsLog", entry_point="r_mabm_b200.environments:CatsLog")
except Exception:
    pass

__all__ = ["Cats", "CatsLog"]
