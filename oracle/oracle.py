"""ctypes wrapper around the CPU oracle (oracle/cats_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under r_mabm_b200/ imports it.

PARITY UNPINNED: see the header of cats_oracle.c -- the reference's arithmetic (ABCredit.jl) is
not on disk, and the reference's tests hold no golden vectors for the step.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libcats_oracle.so"

# Cats._default_parameters, /root/reference/rmabm/environments/cats.py:69-104 (key order matters)
PARAM_KEYS = [
    "z_c", "z_k", "z_e", "xi", "chi", "q_adj", "p_adj", "mu", "eta", "Iprob", "phi", "theta",
    "delta", "alpha", "k", "div", "barX", "inventory_depreciation", "b1", "b2", "b_k1", "b_k2",
    "interest_rate", "subsidy", "maastricht", "target_deficit", "tax_rate", "wage_update_up",
    "wage_update_down", "u_target", "wb", "tax_rate_d", "r_f", "E_threshold_scale",
]
DEFAULT_PARAMS = dict(zip(PARAM_KEYS, [
    5, 2, 5, 0.96, 0.05, 0.9, 0.1, 1.2, 0.03, 0.25, 0.002, 0.05, 0.5, 0.66667, 0.33333, 0.2, 0.85,
    0.3, -15, 13, -5, 5, 0.01, 0, 0.0, 0.01, 0.0, 0.1, 0.01, 0.1, 1.0, 0.0, 0.01, 0,
]))

SCAL_NAMES = [
    "timestep", "price", "price_k", "init_price", "init_price_k", "wb", "Y_real", "Y_nominal_tot",
    "gdp_deflator", "inflationRate", "Un", "consumption", "totK", "bankruptcy_rate", "bank_E",
    "bank_loans", "profitsB", "dividendsB", "bank_interest_income", "bank_bad_debt", "gov_bonds",
    "GB", "deficitGDP", "gov_TA", "gov_G", "gov_r_b", "gov_subsidy_level", "terminated", "status",
    "episode", "reserved1", "reserved2",
]
C_FIELDS = ["De", "P", "Y_prev", "Yd", "Q", "A", "K", "investment", "capital_value", "wages",
            "interests", "deb", "liquidity", "barYK", "interest_r", "div_income"]
K_FIELDS = ["De", "P", "Y_prev", "Yd", "Q", "A", "wages", "interests", "deb", "liquidity",
            "interest_r", "div_income", "inv", "Yavail"]

FLAG_BURNIN, FLAG_LOG, FLAG_AVG_PRICE, FLAG_REWARD_RMS, FLAG_STRICT = 1, 2, 4, 8, 16
N_STATS = 16


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc, a second or two)."""
    src = _HERE / "cats_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", str(_HERE), "-B", "CC=gcc"], check=True, env=env,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_uint64,
                                    C.c_uint32]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint,
                                  C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_f64.restype = C.POINTER(C.c_double)
        L.oracle_f64.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.oracle_i32.restype = C.POINTER(C.c_int32)
        L.oracle_i32.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.oracle_tape_bytes.restype = C.c_size_t
        L.oracle_tape_bytes.argtypes = [C.c_int] * 6
        L.oracle_tape_offsets.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_uint64)]
        L.oracle_fill_tape.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_set_stop_phase.argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_bankruptcy_reward.argtypes = [C.c_void_p, C.c_double]
        L.oracle_det_exp.restype = C.c_double
        L.oracle_det_exp.argtypes = [C.c_double]
        L.oracle_det_log.restype = C.c_double
        L.oracle_det_log.argtypes = [C.c_double]
        L.oracle_det_sum.restype = C.c_double
        L.oracle_det_sum.argtypes = [C.c_void_p, C.c_int]
        L.oracle_philox.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
        L.oracle_sample_distinct.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_perm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_run_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.oracle_run_batch_tape.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint,
                                            C.c_void_p, C.c_int]
        L.oracle_obs.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_void_p]
        _lib = L
    return _lib


def params_vector(params: dict | None = None) -> np.ndarray:
    p = dict(DEFAULT_PARAMS)
    if params:
        p.update(params)
    return np.array([float(p[k]) for k in PARAM_KEYS], dtype=np.float64)


class OracleEconomy:
    """One economy of the CPU restatement."""

    def __init__(self, W=1000, F=100, N=20, params=None, seed=0, economy=0,
                 bankruptcy_reward=-100.0):
        self.W, self.F, self.N, self.H = W, F, N, W + F + N
        self.pv = params_vector(params)
        self.z_c, self.z_k, self.z_e = int(self.pv[0]), int(self.pv[1]), int(self.pv[2])
        self._L = lib()
        self._h = self._L.oracle_create(W, F, N, self.pv.ctypes.data_as(C.POINTER(C.c_double)),
                                        C.c_uint64(seed), C.c_uint32(economy))
        self._L.oracle_set_bankruptcy_reward(self._h, float(bankruptcy_reward))

    def __del__(self):
        try:
            if self._h:
                self._L.oracle_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- state access (numpy views onto the C arrays; copy if you need to keep them) ------------
    def f64(self, name: str) -> np.ndarray:
        n = C.c_int(0)
        p = self._L.oracle_f64(self._h, name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def i32(self, name: str) -> np.ndarray:
        n = C.c_int(0)
        p = self._L.oracle_i32(self._h, name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def scal(self, name: str) -> float:
        return float(self.f64("scal")[SCAL_NAMES.index(name)])

    def snapshot(self) -> dict:
        out = {"scal": self.f64("scal").copy(), "PA": self.f64("PA").copy(),
               "perm": self.f64("perm").copy(), "Oc": self.i32("Oc").copy(),
               "c_Leff": self.i32("c_Leff").copy(), "k_Leff": self.i32("k_Leff").copy()}
        for f in C_FIELDS:
            out["c_" + f] = self.f64("c_" + f).copy()
        for f in K_FIELDS:
            out["k_" + f] = self.f64("k_" + f).copy()
        return out

    # -- stepping ------------------------------------------------------------------------------
    def tape_bytes(self) -> int:
        return int(self._L.oracle_tape_bytes(self.W, self.F, self.N, self.z_c, self.z_k, self.z_e))

    def fill_tape(self) -> np.ndarray:
        rec = np.zeros(self.tape_bytes(), dtype=np.uint8)
        self._L.oracle_fill_tape(self._h, rec.ctypes.data)
        return rec

    def reset(self) -> None:
        """A new model in place, episode counter + 1 (cats_reset_masked)."""
        self._L.oracle_reset(self._h)

    def set_stop_phase(self, p: int) -> None:
        self._L.oracle_set_stop_phase(self._h, p)

    def step(self, actions=None, mask=None, flags=FLAG_BURNIN, tape=None):
        n_agents = 0 if actions is None else len(actions)
        act = np.ascontiguousarray(actions, dtype=np.float64) if n_agents else None
        msk = (np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None
               else np.ones(n_agents, dtype=np.uint8)) if n_agents else None
        obs = np.zeros((n_agents, 2), dtype=np.float64)
        rew = np.zeros(n_agents, dtype=np.float64)
        tp = None
        if tape is not None:
            tape = np.ascontiguousarray(tape, dtype=np.uint8)
            tp = tape.ctypes.data
        self._L.oracle_step(self._h, act.ctypes.data if n_agents else None,
                            msk.ctypes.data if n_agents else None, n_agents, flags, tp,
                            obs.ctypes.data if n_agents else None,
                            rew.ctypes.data if n_agents else None)
        return obs, rew

    def obs(self, n_agents: int, flags: int = 0) -> np.ndarray:
        out = np.zeros((n_agents, 2), dtype=np.float64)
        if n_agents:
            self._L.oracle_obs(self._h, n_agents, flags, out.ctypes.data)
        return out


def run_batch(econs: list[OracleEconomy], n_periods: int, n_threads: int = 1,
              want_stats: bool = True, tape: np.ndarray | None = None, fill_tape: bool = False,
              flags: int = 0) -> np.ndarray | None:
    """Advance every economy n_periods heuristic periods (OpenMP over economies).
    tape: uint8 [B, n_periods, tape_bytes]; with fill_tape the records are first written with the
    Philox source's draws and then replayed, otherwise they are the caller's (a foreign generator).
    flags: extra step flags (FLAG_STRICT)."""
    L = lib()
    B = len(econs)
    arr = (C.c_void_p * B)(*[e._h for e in econs])
    stats = np.zeros((B, n_periods, N_STATS), dtype=np.float64) if want_stats else None
    sp = stats.ctypes.data if want_stats else None
    if tape is None and not flags:
        L.oracle_run_batch(arr, B, n_periods, sp, n_threads)
    else:
        if tape is not None:
            assert tape.dtype == np.uint8 and tape.flags.c_contiguous
            assert tape.shape == (B, n_periods, econs[0].tape_bytes()), tape.shape
        L.oracle_run_batch_tape(arr, B, n_periods, tape.ctypes.data if tape is not None else None,
                                int(fill_tape), flags, sp, n_threads)
    return stats


def tape_offsets(W, F, N, z_c, z_k, z_e) -> dict:
    out = (C.c_uint64 * 11)()
    lib().oracle_tape_offsets(W, F, N, z_c, z_k, z_e, out)
    names = ["u_price_c", "u_price_k", "u_invest", "fire_key", "job_order", "job_cand",
             "cap_order", "cap_cand", "goods_order", "goods_cand", "bytes"]
    return dict(zip(names, [int(v) for v in out]))
