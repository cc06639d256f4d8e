"""The period kernel's source compiled for a CPU emulation of warps and blocks (tests/simt/simt_emu.h).

TEST INFRASTRUCTURE ONLY -- it exists so that the kernel's logic (commit rounds, replay order,
compaction, multi-warp hand-overs) is checked against the oracle in the CPU test-suite, under several
lane schedules.  Nothing under r_mabm_b200/ imports this; the product has no CPU path.

`EmuCats` offers the part of `BatchedCats`' interface the parity helpers use, over numpy blobs.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import subprocess
from pathlib import Path

import numpy as np
import torch

from r_mabm_b200 import _native as nat
from r_mabm_b200.params import load_parameters, params_vector

_HERE = Path(__file__).resolve().parent
_CSRC = _HERE.parent.parent / "r_mabm_b200" / "csrc"
_BUILD = _HERE / "_build"
_SOURCES = [_HERE / "cats_emu.cpp", _HERE / "simt_emu.cpp", _HERE / "simt_emu.h"] + sorted(_CSRC.glob("*.cuh")) + \
    sorted(_CSRC.glob("*.h")) + sorted((_HERE.parent.parent / "include").glob("*.h"))
_lib = None


def build() -> Path:
    import os
    extra = os.environ.get("CATS_EMU_EXTRA", "").split()   # development aid: extra -D flags for the emulated build
    h = hashlib.sha1(" ".join(extra).encode())
    for p in _SOURCES:
        h.update(p.read_bytes())
    out = _BUILD / f"libcats_emu_{h.hexdigest()[:12]}.so"
    if not out.exists():
        _BUILD.mkdir(exist_ok=True)
        if not extra:
            for old in _BUILD.glob("libcats_emu_*.so"):
                old.unlink()
        subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w", *extra,
                        "-I/usr/local/cuda/include", "-o", str(out), str(_HERE / "cats_emu.cpp"),
                        str(_HERE / "simt_emu.cpp")], check=True)
    return out


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(str(build()))
        L.emu_init.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]
        L.emu_step.argtypes = [C.POINTER(nat.StepArgs), C.c_int, C.c_int]
        L.emu_set_schedule.argtypes = [C.c_int]
        L.emu_step_mw.argtypes = [C.POINTER(nat.StepArgs), C.c_int, C.c_int, C.c_int]
        _lib = L
    return _lib


class EmuCats:
    """B economies stepped by the emulated kernel.  mode: 0 tape/debug kernel, 1 production, 2 the
    default-configuration kernel; g: economies per block (1 or 2); schedule: 0 lanes run in order,
    1 in reverse, >= 2 a pseudo-random order per scheduling round."""

    def __init__(self, B, W=1000, F=100, N=20, params=None, n_agents=0, seed=0, economy_base=0, mode=1, g=1,
                 schedule=0, strict=False, flags=0, bankruptcy_reward=-100.0, warps=1, pipe=False):
        self.warps = warps      # > 1: the multi-warp kernel (one block per economy)
        self.pipe = pipe        # its goods market as the producer / consumer pipeline instead of the block-wide one
        self.B, self.W, self.F, self.N = B, W, F, N
        self.n_agents, self.seed, self.economy_base = n_agents, seed, economy_base
        self.mode, self.g, self.schedule, self.strict = mode, g, schedule, strict
        self.base_flags, self.bankruptcy_reward = flags, bankruptcy_reward
        if isinstance(params, (list, tuple)):
            dicts = [load_parameters(p) for p in params]
            self._params = np.array([params_vector(d) for d in dicts], dtype=np.float64)
            self._stride = nat.N_PARAMS
            zs = {(int(d["z_c"]), int(d["z_k"]), int(d["z_e"])) for d in dicts}
            self._z = next(iter(zs)) if len(zs) == 1 else None
            self._max_z = max(max(z) for z in zs)
        else:
            d = load_parameters(params)
            self._params = np.array(params_vector(d), dtype=np.float64)
            self._stride = 0
            self._z = (int(d["z_c"]), int(d["z_k"]), int(d["z_e"]))
            self._max_z = max(self._z)
        self.blob_bytes = int(nat.load().cats_blob_bytes(W, F, N))
        raw = np.zeros(B * self.blob_bytes + 128, dtype=np.uint8)
        o = (-raw.ctypes.data) % 128
        self._blob = raw[o:o + B * self.blob_bytes].reshape(B, self.blob_bytes)
        self._t = torch.from_numpy(self._blob)
        nA = max(n_agents, 1)
        self.obs = np.zeros((B, nA, 2)); self.reward = np.zeros((B, nA))
        self.terminated = np.zeros(B, dtype=np.uint8); self.status = np.zeros(B, dtype=np.uint32)
        lib().emu_init(self._blob.ctypes.data, B, W, F, N, self._params.ctypes.data, self._stride, None)

    # -- the BatchedCats surface parity_util uses ----------------------------------------------
    def field(self, name):
        off, cnt, dt = nat.field_info(self.W, self.F, self.N, name)
        st = nat.field_stride(name)
        v = self._t.view(torch.float64 if dt == 0 else torch.int32)
        e = 8 if dt == 0 else 4
        return v[:, off // e: off // e + cnt * st: st]

    @property
    def blobs(self):
        return self._t

    def workset_bytes(self):
        return int(nat.load().cats_workset_bytes(self.W, self.F, self.N))

    def tape_bytes(self):
        return int(nat.load().cats_tape_bytes(self.W, self.F, self.N, *self._z))

    def reset_where(self, active):
        active = np.ascontiguousarray(active, dtype=np.uint8)
        lib().emu_init(self._blob.ctypes.data, self.B, self.W, self.F, self.N, self._params.ctypes.data, self._stride,
                       active.ctypes.data)

    def _call(self, n_periods, flags, actions=None, mask=None, stats=None, tape=None, stop_phase=0, debug=None,
              active=None):
        a = nat.StepArgs()
        a.d_blobs = self._blob.ctypes.data; a.B = self.B
        a.W, a.F, a.N = self.W, self.F, self.N
        a.d_params = self._params.ctypes.data; a.params_stride = self._stride
        a.seed = self.seed; a.economy_base = self.economy_base
        a.n_periods = n_periods
        a.flags = flags | self.base_flags | (nat.FLAG_STRICT if self.strict else 0)
        a.n_agents = self.n_agents
        keep = [actions, mask, stats, tape, debug, active]
        if actions is not None:
            a.d_actions = actions.ctypes.data
        if mask is not None:
            a.d_action_mask = mask.ctypes.data
        a.bankruptcy_reward = self.bankruptcy_reward
        if self.n_agents > 0:
            a.d_obs = self.obs.ctypes.data; a.d_reward = self.reward.ctypes.data
        a.d_terminated = self.terminated.ctypes.data; a.d_status = self.status.ctypes.data
        if stats is not None:
            a.d_stats = stats.ctypes.data
        if tape is not None:
            a.d_tape = tape.ctypes.data
            a.tape_z[0], a.tape_z[1], a.tape_z[2] = self._z
        a.max_z = self._max_z
        if self._z is not None:
            a.z_all[0], a.z_all[1], a.z_all[2] = self._z
        a.stop_phase = stop_phase
        if debug is not None:
            a.d_debug = debug.ctypes.data
        if active is not None:
            a.d_active = active.ctypes.data
        mode = 0 if (tape is not None or stop_phase or debug is not None) else self.mode
        lib().emu_set_schedule(self.schedule)
        if self.warps > 1 and mode != 0:
            rc = lib().emu_step_mw(C.byref(a), mode, self.warps, int(self.pipe))
        else:
            rc = lib().emu_step(C.byref(a), mode, self.g)
        assert rc == 0, "emulated kernel: unsupported combination"
        del keep

    def run(self, n_periods, want_stats=False, tape=None, active=None):
        stats = np.zeros((self.B, n_periods, nat.N_STATS)) if want_stats else None
        if tape is not None:
            tape = np.ascontiguousarray(tape, dtype=np.uint8)
        if active is not None:
            active = np.ascontiguousarray(active, dtype=np.uint8)
        self._call(n_periods, nat.FLAG_BURNIN, stats=stats, tape=tape, active=active)
        return stats

    def step(self, actions=None, mask=None, burnin=False, want_stats=False, tape=None, stop_phase=0, debug=None,
             active=None):
        stats = np.zeros((self.B, 1, nat.N_STATS)) if want_stats else None
        if actions is not None:
            actions = np.ascontiguousarray(actions, dtype=np.float64)
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
        if tape is not None:
            tape = np.ascontiguousarray(tape, dtype=np.uint8)
        if active is not None:
            active = np.ascontiguousarray(active, dtype=np.uint8)
        self._call(1, nat.FLAG_BURNIN if burnin else 0, actions=actions, mask=mask, stats=stats, tape=tape,
                   stop_phase=stop_phase, debug=debug, active=active)
        return self.obs[:, :self.n_agents], self.reward[:, :self.n_agents], self.terminated, stats
