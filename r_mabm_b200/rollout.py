"""Device-resident rollouts: thousands of economies, learned firm agents, zero host copies.

`BatchedQPolicy` is the torch restatement of the reference's tabular `QLearner`
(/root/reference/rmabm/agents/qlearners.py:63-166) with one table per (economy, agent):
nearest-pole binning, epsilon-greedy on an n2 x n3 action grid, TD(0) update.  `BatchedRollout`
is the batched twin of `Simulation.step` / `run_episode` (/root/reference/rmabm/simulation.py:93-149)
for BASELINE.json's MARL configuration (`examples/agent_training.py`: CatsLog, 2 learned firms): the
policy's action tensor feeds the period kernel directly, observations and rewards never leave the GPU.

`shard_range` / `gather_stats` are the multi-GPU plumbing: economies are independent, so ranks own
contiguous slabs of global economy ids and the only collective is the final gather of statistics.
"""
from __future__ import annotations

import torch

from .engine import BatchedCats


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slab [begin, end) of `total` economies owned by `rank` (remainder to low ranks)."""
    if not 0 <= rank < world:
        raise ValueError("rank must be in [0, world)")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_stats(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather per-economy statistics [B_local, ...] along dim 0 (NCCL on GPUs, gloo in CPU tests).
    Slabs may differ in size by one economy; they are padded for the collective and trimmed after."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], 0)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)


class BatchedQPolicy:
    """Independent tabular Q-learners, one per (economy, agent), as one [B, A, n0, n1, n2, n3] tensor."""

    def __init__(self, B: int, n_agents: int, bounds: dict, n_bins=(11, 11, 11, 11), alpha: float = 0.1,
                 gamma: float = 0.99, epsilon: float = 0.9, device="cuda", seed: int = 0):
        if isinstance(n_bins, int):
            n_bins = (n_bins,) * 4
        self.B, self.A, self.n = B, n_agents, tuple(n_bins)
        self.alpha, self.gamma, self.epsilon = alpha, gamma, epsilon
        self.bounds = bounds
        self.device = torch.device(device)
        self.Q = torch.zeros((B, n_agents) + self.n, dtype=torch.float64, device=self.device)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)
        self._last = None  # (s0, s1, a0, a1) of the last get_action
        self._b = torch.arange(B, device=self.device).view(B, 1).expand(B, n_agents)
        self._a = torch.arange(n_agents, device=self.device).view(1, n_agents).expand(B, n_agents)

    def _bin(self, x: torch.Tensor, key: str, n: int) -> torch.Tensor:
        """Index of the nearest of n poles linearly spaced over bounds[key]; the lower pole wins a tie,
        out-of-range values clamp (qlearners.py:83-94)."""
        lo, hi = self.bounds[key]
        step = (hi - lo) / (n - 1)
        idx = torch.ceil((x - lo) / step - 0.5)
        idx = torch.nan_to_num(idx, nan=0.0, posinf=float(n - 1), neginf=0.0)
        return idx.clamp_(0, n - 1).to(torch.int64)

    def bin_obs(self, obs: torch.Tensor):
        return self._bin(obs[..., 0], "obs_firm_stock", self.n[0]), self._bin(obs[..., 1], "obs_price_delta", self.n[1])

    def get_action(self, obs: torch.Tensor) -> torch.Tensor:
        """obs [B, A, 2] -> actions [B, A, 2] (production factor, price factor), all on device."""
        s0, s1 = self.bin_obs(obs)
        q = self.Q[self._b, self._a, s0, s1].reshape(self.B, self.A, -1)
        greedy = q.argmax(-1)
        g0, g1 = greedy // self.n[3], greedy % self.n[3]
        explore = torch.rand((self.B, self.A), generator=self.gen, device=self.device) < self.epsilon
        r0 = torch.randint(0, self.n[2], (self.B, self.A), generator=self.gen, device=self.device)
        r1 = torch.randint(0, self.n[3], (self.B, self.A), generator=self.gen, device=self.device)
        a0 = torch.where(explore, r0, g0)
        a1 = torch.where(explore, r1, g1)
        self._last = (s0, s1, a0, a1)
        lo0, hi0 = self.bounds["act_production_factor"]
        lo1, hi1 = self.bounds["act_price_factor"]
        act = torch.stack([lo0 + a0.to(torch.float64) * (hi0 - lo0) / (self.n[2] - 1),
                           lo1 + a1.to(torch.float64) * (hi1 - lo1) / (self.n[3] - 1)], -1)
        return act.contiguous()

    def train(self, reward: torch.Tensor, next_obs: torch.Tensor, terminated: torch.Tensor) -> None:
        """Q[s,a] += alpha (r + gamma max Q[s'] - Q[s,a]); no bootstrap where terminated (qlearners.py:156-166)."""
        s0, s1, a0, a1 = self._last
        n0, n1 = self.bin_obs(next_obs)
        boot = self.Q[self._b, self._a, n0, n1].reshape(self.B, self.A, -1).max(-1).values
        term = terminated.to(torch.bool).view(self.B, 1)
        target = reward + torch.where(term, torch.zeros_like(boot), self.gamma * boot)
        cur = self.Q[self._b, self._a, s0, s1, a0, a1]
        self.Q[self._b, self._a, s0, s1, a0, a1] = cur + self.alpha * (target - cur)


class FusedQPolicy(BatchedQPolicy):
    """The same learners with `train` and `get_action` as ONE hand-written kernel per rollout step
    (`cats_q_step`, include/cats_b200.h) instead of ~45 torch launches: the rollout of BASELINE.json
    config 4 is launch-bound otherwise.  Same binning, TD(0) update and greedy rule as `BatchedQPolicy`
    (qlearners.py:63-166).  Exploration draws come from Philox keyed by (seed, global economy id, agent,
    step) -- not from a torch generator -- so a rollout is reproducible at any batch size or sharding.
    CUDA only; `BatchedQPolicy` is the torch twin that also runs on the CPU."""

    def __init__(self, B: int, n_agents: int, bounds: dict, n_bins=(11, 11, 11, 11), alpha: float = 0.1,
                 gamma: float = 0.99, epsilon: float = 0.9, device="cuda", seed: int = 0, economy_base: int = 0):
        super().__init__(B, n_agents, bounds, n_bins, alpha, gamma, epsilon, device, seed)
        if self.device.type != "cuda":
            raise RuntimeError("FusedQPolicy runs on a CUDA device; use BatchedQPolicy on the CPU")
        from . import _native as nat
        self._nat, self._lib = nat, nat.load()
        self.seed, self.economy_base, self.step_count = int(seed), int(economy_base), 0
        self.last = torch.zeros((B, n_agents, 4), dtype=torch.int32, device=self.device)
        # two action buffers, written alternately: the actions of the step in flight stay readable (loggers,
        # tests) while the next ones are being chosen, without a copy
        self._act_bufs = [torch.zeros((B, n_agents, 2), dtype=torch.float64, device=self.device) for _ in range(2)]
        self.actions = self._act_bufs[0]

    def _launch(self, obs, reward, terminated, do_train: bool, do_act: bool, active=None) -> None:
        import ctypes as C
        a = self._nat.QArgs()
        a.d_Q = self.Q.data_ptr(); a.B = self.B; a.A = self.A
        for i in range(4):
            a.n[i] = self.n[i]
        b = self.bounds
        a.obs_lo[0], a.obs_hi[0] = b["obs_firm_stock"]; a.obs_lo[1], a.obs_hi[1] = b["obs_price_delta"]
        a.act_lo[0], a.act_hi[0] = b["act_production_factor"]; a.act_lo[1], a.act_hi[1] = b["act_price_factor"]
        a.alpha, a.gamma, a.epsilon = self.alpha, self.gamma, self.epsilon
        a.seed = self.seed & 0xFFFFFFFFFFFFFFFF; a.economy_base = self.economy_base
        a.step = self.step_count & 0xFFFFFFFF
        a.do_train, a.do_act = int(do_train), int(do_act)
        for t in (obs, reward, terminated):
            if t is not None and not (t.is_cuda and t.is_contiguous()):
                raise ValueError("FusedQPolicy takes contiguous CUDA tensors")
        a.d_obs = obs.data_ptr()
        a.d_reward = reward.data_ptr() if reward is not None else None
        a.d_terminated = terminated.data_ptr() if terminated is not None else None
        if do_act:
            self.actions = self._act_bufs[self.step_count & 1]
        a.d_last = self.last.data_ptr(); a.d_actions = self.actions.data_ptr()
        a.d_active = active.data_ptr() if active is not None else None
        a.stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            self._nat.check(self._lib.cats_q_step(C.byref(a)))

    def get_action(self, obs: torch.Tensor, active: torch.Tensor | None = None) -> torch.Tensor:
        self._launch(obs, None, None, False, True, active)
        self.step_count += 1
        return self.actions

    def train(self, reward: torch.Tensor, next_obs: torch.Tensor, terminated: torch.Tensor) -> None:
        self._launch(next_obs, reward, terminated, True, False)

    def train_and_act(self, reward: torch.Tensor, next_obs: torch.Tensor, terminated: torch.Tensor) -> torch.Tensor:
        """TD update for the transition that ended in `next_obs`, then the action for it: one launch."""
        self._launch(next_obs, reward, terminated, True, True)
        self.step_count += 1
        return self.actions


class BatchedRollout:
    """env step + policy + TD update on device for B economies (BASELINE.json config 4).

    `auto_reset=True` gives every economy its own episode clock, as `Simulation.run_episode`
    (simulation.py:123-149) does for the one economy of the reference: an economy that is terminated or
    has run its T steps gets a new model and its burn-in (`BatchedCats.reset_where`) while the others
    carry on, all without a host synchronisation; its next observation is the one `Cats.reset` would
    return.  `episodes` counts finished episodes per economy."""

    def __init__(self, env: BatchedCats, policy: BatchedQPolicy, t_burnin: int = 300, T: int = 1000,
                 auto_reset: bool = False):
        self.env, self.policy, self.t_burnin, self.T = env, policy, t_burnin, T
        self.auto_reset = auto_reset
        self.t = 0
        self.obs = None
        self._next_actions = None
        self.t_b = torch.zeros(env.B, dtype=torch.int32, device=env.device)        # steps into the episode
        self.episodes = torch.zeros(env.B, dtype=torch.int32, device=env.device)   # finished episodes

    def _burnin(self, active: torch.Tensor | None = None) -> None:
        """t_burnin heuristic periods (cats.py:423-432); the last one is launched with outputs, so
        the observation buffer holds what `Cats.reset` returns (`_get_obs` of the burnt-in state)."""
        if self.t_burnin > 1:
            self.env.run(self.t_burnin - 1, active=active)
        if self.t_burnin > 0:
            self.env.step(burnin=True, active=active)
        else:
            self.env.observe(active=active)

    def reset(self, seed: int | None = None) -> torch.Tensor:
        self.env.reset_state(seed)
        self._burnin()
        self.t = 0
        self._next_actions = None
        self.t_b.zero_(); self.episodes.zero_()
        self.obs = self.env.observation()
        return self.obs

    def step(self, train: bool = True):
        fused = train and not self.auto_reset and hasattr(self.policy, "train_and_act")
        if fused and self._next_actions is not None:
            actions = self._next_actions          # chosen by the launch that trained on the last transition
        else:
            # an action cached by an earlier fused step was chosen for an older observation: drop it
            self._next_actions = None
            actions = self.policy.get_action(self.obs)
        self.last_actions = actions      # kept for loggers / tests; never leaves the device
        obs, reward, terminated, _ = self.env.step(actions)
        if fused:
            # two launches per rollout step: the period kernel, then TD update + next action in one
            # (`actions` stays valid: the policy writes the next ones into its other buffer)
            self._next_actions = self.policy.train_and_act(reward, obs, terminated)
        elif train:
            self.policy.train(reward, obs, terminated)
        self.t += 1
        if self.auto_reset:
            reward, terminated = reward.clone(), terminated.clone()   # the burn-in below reuses the buffers
            self.t_b += 1
            done = terminated.to(torch.bool) | (self.t_b >= self.T)
            self.last_done = done
            self.env.reset_where(done)
            self._burnin(active=done)
            self.t_b.masked_fill_(done, 0)
            self.episodes += done.to(torch.int32)
            obs = self.env.observation()
        # a view of the engine's observation buffer: valid until the next step overwrites it
        self.obs = obs
        return reward, terminated

    def run(self, n_steps: int, train: bool = True) -> torch.Tensor:
        """Returns the per-step mean reward [n_steps] (device tensor)."""
        if self.obs is None:
            self.reset()
        out = torch.empty(n_steps, dtype=torch.float64, device=self.env.device)
        for i in range(n_steps):
            r, _ = self.step(train)
            out[i] = r.mean()
        return out
