"""BatchedCats: B independent CATS economies resident on one B200, stepped by libcats_b200.so.

This is the batched core behind the reference-shaped `Cats` / `CatsLog` adapters
(r_mabm_b200/environments/cats.py).  It owns one torch uint8 tensor holding the B economy blobs
and hands raw device pointers to the C-ABI (include/cats_b200.h).  PyTorch is plumbing here:
device memory, streams, and (in rollout.py / bench.py) torch.distributed.

Replaces, per economy: `ABCredit.initialise_model` and the 26 ABCredit calls of `Cats.step`
(/root/reference/rmabm/environments/cats.py:224, 355-413) plus the Python-side action override,
reward, termination and observation code (cats.py:434-556).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict

import torch

from . import _native as nat
from .params import load_parameters, params_vector, validate


class BatchedCats:
    """B economies of size (W, F, N) on one CUDA device.

    Args mirror `Cats.__init__` (cats.py:113-128) where they apply to the model; the gym-only
    arguments live in the adapters.  `params` may be one dict (shared) or a list of B dicts
    (heterogeneous economies).  `economy_base` is the global id of economy 0 of this shard.
    """

    def __init__(self, B: int, W: int = 1000, F: int = 100, N: int = 20,
                 params: Dict[str, Any] | list | None = None, n_agents: int = 0,
                 reward_type: str = "profits", price_change: str = "agent_price",
                 log_mode: bool = False, bankruptcy_reward: float = -100.0, seed: int = 0,
                 economy_base: int = 0, device: str | torch.device = "cuda", strict: bool = False):
        # strict: the plain arithmetic of docs/MODEL.md 4.6 (CATS_FLAG_STRICT) -- divisions by the price and
        # an fp64 demand sum in visiting order instead of reciprocals and the order-free fixed-point sum
        self.strict = bool(strict)
        if reward_type not in ("profits", "rms"):
            raise ValueError("Invalid reward type. Must be 'profits' or 'rms'")
        if price_change not in ("agent_price", "avg_price"):
            raise ValueError("Invalid price change mechanism. Must be 'agent_price' or 'avg_price'")
        if B < 1:
            raise ValueError("B must be >= 1")
        if not 0 <= n_agents <= F:
            raise ValueError("n_agents must be in [0, F]")
        self._lib = nat.load()
        if not torch.cuda.is_available():
            raise RuntimeError("r_mabm_b200 needs a CUDA device: there is no CPU fallback")
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.B, self.W, self.F, self.N, self.H = B, W, F, N, W + F + N
        self.n_agents = n_agents
        self.reward_type, self.price_change, self.log_mode = reward_type, price_change, log_mode
        self.bankruptcy_reward = float(bankruptcy_reward)
        self.seed, self.economy_base = int(seed), int(economy_base)

        if isinstance(params, (list, tuple)):
            if len(params) != B:
                raise ValueError("a list of params must have one entry per economy")
            dicts = [load_parameters(p) for p in params]
            for d in dicts:
                validate(d, F, N)
            zs = {(int(d["z_c"]), int(d["z_k"]), int(d["z_e"])) for d in dicts}
            self.params = dicts[0]
            self._z = next(iter(zs)) if len(zs) == 1 else None
            self._params_t = torch.tensor([params_vector(d) for d in dicts], dtype=torch.float64,
                                          device=self.device)
            self._params_stride = nat.N_PARAMS
            self._max_z = max(max(int(d[k]) for k in ("z_c", "z_k", "z_e")) for d in dicts)
        else:
            self.params = load_parameters(params)
            validate(self.params, F, N)
            self._z = (int(self.params["z_c"]), int(self.params["z_k"]), int(self.params["z_e"]))
            self._params_t = torch.tensor(params_vector(self.params), dtype=torch.float64,
                                          device=self.device)
            self._params_stride = 0
            self._max_z = max(self._z)

        self.blob_bytes = int(self._lib.cats_blob_bytes(W, F, N))
        if self.blob_bytes == 0:
            nat.check(nat.EINVAL)
        with torch.cuda.device(self.device):
            # torch's caching allocator returns 512-byte aligned blocks; blobs are multiples of 128 B
            self._blob = torch.zeros((B, self.blob_bytes), dtype=torch.uint8, device=self.device)
            assert self._blob.data_ptr() % 128 == 0
            nA = max(n_agents, 1)
            self._obs = torch.zeros((B, nA, 2), dtype=torch.float64, device=self.device)
            self._reward = torch.zeros((B, nA), dtype=torch.float64, device=self.device)
            self._terminated = torch.zeros((B,), dtype=torch.uint8, device=self.device)
            self._status = torch.zeros((B,), dtype=torch.int32, device=self.device)
        self._f64 = self._blob.view(torch.float64)
        self._i32 = self._blob.view(torch.int32)
        self._views: dict[str, torch.Tensor] = {}
        self._host_scratch = None
        self._host_call_cache: dict = {}   # step_host: argument blocks by call shape
        self.periods_done = 0
        self.reset_state()

    # -- state views -------------------------------------------------------------------------
    def field(self, name: str) -> torch.Tensor:
        """Zero-copy [B, count] view of a named state field (see cats_field_info)."""
        v = self._views.get(name)
        if v is None:
            off, cnt, dt = nat.field_info(self.W, self.F, self.N, name)
            st = nat.field_stride(name)        # 2 for the {PA, perm} household records
            v = (self._f64[:, off // 8: off // 8 + cnt * st: st] if dt == 0
                 else self._i32[:, off // 4: off // 4 + cnt * st: st])
            self._views[name] = v
        return v

    def scalar(self, name: str) -> torch.Tensor:
        return self.field("scal")[:, SCAL_INDEX[name]]

    @property
    def blobs(self) -> torch.Tensor:
        return self._blob

    @property
    def flags(self) -> int:
        f = 0
        if self.log_mode:
            f |= nat.FLAG_LOG
        if self.price_change == "avg_price":
            f |= nat.FLAG_AVG_PRICE
        if self.reward_type == "rms":
            f |= nat.FLAG_REWARD_RMS
        if self.strict:
            f |= nat.FLAG_STRICT
        return f

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # -- model life cycle --------------------------------------------------------------------
    def reset_state(self, seed: int | None = None) -> None:
        """`initialise_model` for every economy (cats.py:224).  No burn-in here."""
        if seed is not None:
            self.seed = int(seed)
        with torch.cuda.device(self.device):
            nat.check(self._lib.cats_init(self._blob.data_ptr(), self.B, self.W, self.F, self.N,
                                          self._params_t.data_ptr(), self._params_stride,
                                          self._stream()))
        self.periods_done = 0

    def _active(self, active: torch.Tensor | None) -> torch.Tensor | None:
        if active is None:
            return None
        if tuple(active.shape) != (self.B,) or not active.is_cuda:
            raise ValueError(f"active must be a CUDA tensor of shape ({self.B},)")
        return active.to(torch.uint8).contiguous() if active.dtype != torch.uint8 else active.contiguous()

    def reset_where(self, active: torch.Tensor, t_burnin: int = 0) -> None:
        """Per-economy `Cats.reset` (cats.py:283-315) without leaving the device: the economies whose
        `active[b]` is set get a new model and `t_burnin` heuristic periods; the others are not
        touched.  A reset economy's episode counter goes up by one and salts its Philox stream, so the
        new episode is a fresh draw that still depends only on (seed, economy id, episode, period).
        `active` is a bool / uint8 CUDA tensor [B] -- e.g. `terminated | truncated` -- and is never
        read on the host: no synchronisation, two launches."""
        act = self._active(active)
        with torch.cuda.device(self.device):
            nat.check(self._lib.cats_reset_masked(self._blob.data_ptr(), self.B, self.W, self.F, self.N,
                                                  self._params_t.data_ptr(), self._params_stride,
                                                  act.data_ptr(), self._stream()))
        if t_burnin > 0:
            self.run(t_burnin, active=act)

    def _args(self, n_periods: int, flags: int, actions=None, mask=None, stats=None, tape=None,
              stop_phase: int = 0, debug=None, want_io: bool = True, active=None) -> nat.StepArgs:
        if active is not None and (tape is not None or stop_phase or debug is not None):
            raise ValueError("`active` cannot be combined with a replay tape, stop_phase or a debug image")
        a = nat.StepArgs()
        a.d_blobs = self._blob.data_ptr(); a.B = self.B
        a.W, a.F, a.N = self.W, self.F, self.N
        a.d_params = self._params_t.data_ptr(); a.params_stride = self._params_stride
        a.seed = self.seed & 0xFFFFFFFFFFFFFFFF; a.economy_base = self.economy_base
        a.n_periods = n_periods; a.flags = flags; a.n_agents = self.n_agents
        a.d_actions = actions.data_ptr() if actions is not None else None
        a.d_action_mask = mask.data_ptr() if mask is not None else None
        a.bankruptcy_reward = self.bankruptcy_reward
        if want_io and self.n_agents > 0:
            a.d_obs = self._obs.data_ptr(); a.d_reward = self._reward.data_ptr()
        a.d_terminated = self._terminated.data_ptr(); a.d_status = self._status.data_ptr()
        a.d_stats = stats.data_ptr() if stats is not None else None
        if tape is not None:
            if self._z is None:
                raise ValueError("tape replay needs the same z_c, z_k, z_e in every economy")
            a.d_tape = tape.data_ptr()
            a.tape_z[0], a.tape_z[1], a.tape_z[2] = self._z
        a.max_z = self._max_z
        if self._z is not None:
            a.z_all[0], a.z_all[1], a.z_all[2] = self._z
        a.stop_phase = stop_phase
        a.d_debug = debug.data_ptr() if debug is not None else None
        a.stream = self._stream()
        a.d_active = active.data_ptr() if active is not None else None
        return a

    def run(self, n_periods: int, want_stats: bool = False, tape: torch.Tensor | None = None,
            active: torch.Tensor | None = None):
        """Advance every economy `n_periods` heuristic periods in ONE launch (state stays on chip
        between periods).  This is the burn-in of `Cats.reset` (cats.py:423-432) and the
        Monte-Carlo path.  Returns [B, n_periods, 16] statistics if asked.  `active` ([B] bool/uint8
        CUDA tensor): only those economies advance (rows of the others in the statistics are zero)."""
        stats = None
        active = self._active(active)
        with torch.cuda.device(self.device):
            if want_stats:
                stats = (torch.empty if active is None else torch.zeros)(
                    (self.B, n_periods, nat.N_STATS), dtype=torch.float64, device=self.device)
            if tape is not None:
                self._check_tape(tape, n_periods)
            a = self._args(n_periods, self.flags | nat.FLAG_BURNIN, stats=stats, tape=tape,
                           want_io=False, active=active)
            nat.check(self._lib.cats_step(C.byref(a)))
        self.periods_done += n_periods
        return stats

    def step(self, actions: torch.Tensor | None = None, mask: torch.Tensor | None = None,
             burnin: bool = False, want_stats: bool = False, tape: torch.Tensor | None = None,
             stop_phase: int = 0, debug: torch.Tensor | None = None, active: torch.Tensor | None = None):
        """One period for every economy with RL actions (device tensors, zero copies).
        `active` ([B] bool/uint8 CUDA tensor): only those economies advance; the observation, reward
        and terminated entries of the others keep their previous values.

        actions: float64 [B, n_agents, 2] (production factor, price factor); mask: uint8/bool
        [B, n_agents], False = "dummy" action (cats.py:439-440).  Returns
        (obs [B,A,2], reward [B,A], terminated [B] uint8, stats [B,1,16] or None); the first three
        are views of buffers that the next step overwrites.
        """
        flags = self.flags | (nat.FLAG_BURNIN if burnin else 0)
        if self.n_agents > 0 and not burnin:
            if actions is None:
                raise ValueError(f"Number of actions (0) must be equal to the number of agents ({self.n_agents})")
            if tuple(actions.shape) != (self.B, self.n_agents, 2):
                raise ValueError(f"actions must have shape {(self.B, self.n_agents, 2)}, got {tuple(actions.shape)}")
            if actions.dtype != torch.float64 or not actions.is_cuda or not actions.is_contiguous():
                raise ValueError("actions must be a contiguous float64 CUDA tensor")
            if mask is not None:
                if tuple(mask.shape) != (self.B, self.n_agents):
                    raise ValueError("mask must have shape [B, n_agents]")
                mask = mask.to(torch.uint8).contiguous() if mask.dtype != torch.uint8 else mask.contiguous()
        else:
            actions, mask = None, None
        stats = None
        active = self._active(active)
        with torch.cuda.device(self.device):
            if want_stats:
                stats = (torch.empty if active is None else torch.zeros)(
                    (self.B, 1, nat.N_STATS), dtype=torch.float64, device=self.device)
            if tape is not None:
                self._check_tape(tape, 1)
            a = self._args(1, flags, actions=actions, mask=mask, stats=stats, tape=tape,
                           stop_phase=stop_phase, debug=debug, active=active)
            nat.check(self._lib.cats_step(C.byref(a)))
        self.periods_done += 1
        nA = self.n_agents
        return self._obs[:, :nA], self._reward[:, :nA], self._terminated, stats

    def observation(self) -> torch.Tensor:
        """The [B, n_agents, 2] observation buffer the last `step` / `observe` filled (a view)."""
        return self._obs[:, :self.n_agents]

    def observe(self, active: torch.Tensor | None = None) -> torch.Tensor:
        """`_get_obs` (cats.py:466-481, 655-671) from the current state without advancing it -- only
        needed when there is no burn-in step to take the observation from.  Torch arithmetic on the
        state views: exact in linear mode, `torch.log` in log mode."""
        nA = self.n_agents
        yp, yd, p = (self.field(k)[:, :nA] for k in ("c_Y_prev", "c_Yd", "c_P"))
        ap = self.scalar("price").unsqueeze(1)
        if self.log_mode:
            new = torch.stack([torch.log(yp + 1e-8) - torch.log(yd + 1e-8),
                               torch.log(p + 1e-8) - torch.log(ap + 1e-8)], -1)
        else:
            new = torch.stack([yp - yd, p - ap], -1)
        if active is None:
            self._obs[:, :nA] = new
        else:
            m = self._active(active).to(torch.bool).view(self.B, 1, 1)
            self._obs[:, :nA] = torch.where(m, new, self._obs[:, :nA])
        return self._obs[:, :nA]

    def _ensure_host_block(self) -> None:
        nA, B = self.n_agents, self.B
        need = int(self._lib.cats_host_scratch_bytes(B, nA, 1))
        if self._host_scratch is None or self._host_scratch.numel() < need:
            self._host_scratch = torch.empty((need,), dtype=torch.uint8, device=self.device)
            # ONE pinned block laid out like the device staging block: the library then moves the inputs
            # with one copy and the outputs with one copy (cats_host_block_offsets)
            off = (C.c_uint64 * 7)()
            nat.check(self._lib.cats_host_block_offsets(B, nA, 1, off))
            off = [int(x) for x in off]
            blk = self._host_block = torch.empty((need,), dtype=torch.uint8).pin_memory()
            A1 = max(nA, 1)

            def view(o, nbytes, dtype, shape):
                return blk[o:o + nbytes].view(dtype).view(shape)
            self._h_act_in = view(off[0], B * nA * 16, torch.float64, (B, nA, 2)) if nA else None
            self._h_mask_in = view(off[1], B * nA, torch.uint8, (B, nA)) if nA else None
            self._h_obs = view(off[2], B * nA * 16, torch.float64, (B, nA, 2)) if nA else torch.empty((B, A1, 2), dtype=torch.float64)
            self._h_rew = view(off[3], B * nA * 8, torch.float64, (B, nA)) if nA else torch.empty((B, A1), dtype=torch.float64)
            self._h_term = view(off[4], B, torch.uint8, (B,))
            self._h_status = view(off[5], B * 4, torch.int32, (B,))
            self._h_stats = view(off[6], B * nat.N_STATS * 8, torch.float64, (B, 1, nat.N_STATS))

    def step_host(self, actions_host: torch.Tensor | None, mask_host: torch.Tensor | None = None,
                  want_stats: bool = True):
        """End-to-end step with HOST buffers: actions in, obs / reward / terminated / stats out, one stream
        sync.  The buffers are this object's pinned block: one copy moves the actions in, and the kernel writes the
        outputs in place over PCIe as each economy finishes (cats_step_host: zero copy for page-locked output
        buffers, staged copies for pageable ones).  This is the
        call bench.py times for `e2e`."""
        nA, B = self.n_agents, self.B
        self._ensure_host_block()
        acting = nA > 0 and actions_host is not None
        if acting:
            if tuple(actions_host.shape) != (B, nA, 2) or actions_host.dtype != torch.float64 or actions_host.is_cuda:
                raise ValueError("actions_host must be a float64 CPU tensor of shape [B, n_agents, 2]")
            # staged through the pinned block (a host-to-host copy of a few hundred KB) unless the caller
            # already wrote into the block's own views (`host_action_buffers`)
            if actions_host.data_ptr() != self._h_act_in.data_ptr():
                self._h_act_in.copy_(actions_host)
            if mask_host is not None and mask_host.data_ptr() != self._h_mask_in.data_ptr():
                self._h_mask_in.copy_(mask_host)
        # The argument block of a call shape is built once: at a few thousand economies per GPU the step is ~150 us and
        # the Python side of the call is what is left to trim.
        flags = self.flags if acting else self.flags | nat.FLAG_BURNIN
        key = (flags, acting, acting and mask_host is not None, want_stats, self._stream(), self._host_scratch.data_ptr(),
               self.seed, self.economy_base, self.bankruptcy_reward)
        hit = self._host_call_cache.get(key)
        if hit is None:
            a = self._args(1, flags, want_io=False)
            if acting:
                a.d_actions = self._h_act_in.data_ptr()
                a.d_action_mask = self._h_mask_in.data_ptr() if mask_host is not None else None
            if nA > 0:      # a burn-in step (no actions) reports observation and reward too, as the reference's does
                a.d_obs = self._h_obs.data_ptr(); a.d_reward = self._h_rew.data_ptr()
            a.d_terminated = self._h_term.data_ptr(); a.d_status = self._h_status.data_ptr()
            a.d_stats = self._h_stats.data_ptr() if want_stats else None
            hit = (a, C.byref(a), self._host_scratch.data_ptr(),
                   (self._h_obs[:, :nA], self._h_rew[:, :nA], self._h_term, self._h_stats if want_stats else None))
            if len(self._host_call_cache) > 16:
                self._host_call_cache.clear()
            self._host_call_cache[key] = hit
        if torch.cuda.current_device() == self.device.index:
            nat.check(self._lib.cats_step_host(hit[1], hit[2]))
        else:
            with torch.cuda.device(self.device):
                nat.check(self._lib.cats_step_host(hit[1], hit[2]))
        self.periods_done += 1
        return hit[3]

    def host_action_buffers(self):
        """Pinned host views [B, n_agents, 2] float64 and [B, n_agents] uint8 inside the staging block:
        fill these and pass them to `step_host` to skip the host-side staging copy."""
        self._ensure_host_block()
        return self._h_act_in, self._h_mask_in

    def h2d_bytes_per_step(self) -> int:
        return self.B * self.n_agents * 2 * 8

    def d2h_bytes_per_step(self, want_stats: bool = True) -> int:
        nA = self.n_agents
        return self.B * (nA * 2 * 8 + nA * 8 + 1 + 4 + (nat.N_STATS * 8 if want_stats else 0))

    # -- replay tape -------------------------------------------------------------------------
    def tape_bytes(self) -> int:
        if self._z is None:
            raise ValueError("tape replay needs the same z_c, z_k, z_e in every economy")
        return int(self._lib.cats_tape_bytes(self.W, self.F, self.N, *self._z))

    def _check_tape(self, tape: torch.Tensor, n_periods: int) -> None:
        want = self.B * n_periods * self.tape_bytes()
        if tape.dtype != torch.uint8 or not tape.is_cuda or not tape.is_contiguous() or tape.numel() != want:
            raise ValueError(f"tape must be a contiguous uint8 CUDA tensor of {want} bytes")
        if tape.data_ptr() % 16:
            raise ValueError("tape must be 16-byte aligned")

    # -- debugging ---------------------------------------------------------------------------
    def workset_bytes(self) -> int:
        return int(self._lib.cats_workset_bytes(self.W, self.F, self.N))

    def launch_info(self) -> dict:
        t, s, b, n = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
        with torch.cuda.device(self.device):
            nat.check(self._lib.cats_launch_info(self.W, self.F, self.N, C.byref(t), C.byref(s),
                                                 C.byref(b), C.byref(n)))
        return {"threads_per_block": t.value, "economies_per_block": t.value // 32, "smem_bytes_per_block": s.value,
                "economies_per_sm": b.value, "num_sms": n.value}

    def launch_shape(self) -> dict:
        """The kernel shape the library picks for THIS batch (see cats_launch_shape)."""
        w, g, p = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        with torch.cuda.device(self.device):
            nat.check(self._lib.cats_launch_shape(self.W, self.F, self.N, self.B, C.byref(w), C.byref(g), C.byref(p)))
        return {"warps_per_economy": w.value, "economies_per_block": g.value,
                "goods_market": ("pipelined" if p.value else "block-wide") if w.value > 1 else "warp-wide"}

    def algo_bytes_per_period(self) -> int:
        """CATS_ALGO_BYTES (docs/DESIGN.md): every state byte read once and written once."""
        return 2 * self.blob_bytes


SCAL_NAMES = [
    "timestep", "price", "price_k", "init_price", "init_price_k", "wb", "Y_real", "Y_nominal_tot",
    "gdp_deflator", "inflationRate", "Un", "consumption", "totK", "bankruptcy_rate", "bank_E",
    "bank_loans", "profitsB", "dividendsB", "bank_interest_income", "bank_bad_debt", "gov_bonds",
    "GB", "deficitGDP", "gov_TA", "gov_G", "gov_r_b", "gov_subsidy_level", "terminated", "status",
    "episode", "reserved1", "reserved2",
]
SCAL_INDEX = {n: i for i, n in enumerate(SCAL_NAMES)}
