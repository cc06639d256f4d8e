#!/usr/bin/env python
"""bench.py -- economy-periods/sec of the batched CATS step on N B200s (one process per GPU).

  python bench.py --gpus 1 --steps 300 --warmup 10                  # config 3 on one GPU (the driver's line)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P bench.py --gpus N --steps K --warmup W        # config 3, 16,384 economies TOTAL over N GPUs
  python bench.py --config 5                                        # BASELINE.json configs[4] (10x economy)
  python bench.py --impl reference --gpus 1 --steps 5 --warmup 1    # CPU arm (oracle port)

Workloads = BASELINE.json `configs` (1-based here as in BASELINE.md 5):
  2  1,024 default-size economies (W=1000, F=100, N=20), heuristic firms        [configs[1]]
  3  16,384 default-size economies, Monte-Carlo seeds -- THE DEFAULT            [configs[2]]
  4  MARL rollout: 4,096 CatsLog economies x 2 learned firms, env step + fused tabular-Q policy on device [configs[3]]
  5  256 economies of 10x size (W=10000, F=1000, N=200)                         [configs[4]]
A *step* is one period of every economy: one launch of the period kernel through the C-ABI (config 4: the
period kernel + the policy kernel).  `--scaling strong` (default): the config's economies are the TOTAL,
sharded over the ranks with `shard_range` -- BASELINE.json: "16,384 ... sharded over 1/2/4/8 B200";
`--scaling weak`: that many economies PER GPU (round 1's line).  Economies are independent: no collective on
the step path; NCCL is used once, after the timed region, to gather the logged statistics
(`ShardedCats.gather`).

L2: a workload whose per-GPU state exceeds the 126 MB L2 needs no flush (config 3 on <= 4 GPUs, config 4);
the others (config 2, config 5, config 3 on 8 GPUs) get a 256 MB buffer written between timed steps, outside the
per-step event pairs.  `config.l2_policy` says which.

One JSON line on stdout (rank 0).  `value` = device-resident inputs, CUDA events per step, max over ranks;
`e2e` = the same step through the public API with HOST buffers inside the timed region: the actions / mask uploaded
from pinned memory with one copy, obs / reward / terminated / status / stats delivered into pinned host memory (the
kernel writes them in place over PCIe, `cats_step_host`; `--e2e-staged`: explicit copies both ways), one stream sync.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

BURNIN = 300
METRIC = "economy-periods/sec (batched CATS step)"
UNIT = "economy-periods/s"
L2_BYTES = 126e6

CONFIGS = {
    2: dict(W=1000, F=100, N=20, economies=1024, n_agents=1, rollout=False,
            what="BASELINE.json configs[1]: 1,024 default-size economies, heuristic firms"),
    3: dict(W=1000, F=100, N=20, economies=16384, n_agents=1, rollout=False,
            what="BASELINE.json configs[2]: 16,384 default-size economies, Monte-Carlo seeds, heuristic firms"),
    4: dict(W=1000, F=100, N=20, economies=4096, n_agents=2, rollout=True,
            what="BASELINE.json configs[3]: MARL rollout, 4,096 CatsLog economies x 2 learned firms (tabular Q, "
                 "examples/agent_training.py), env step + TD update + epsilon-greedy choice on device"),
    5: dict(W=10000, F=1000, N=200, economies=256, n_agents=1, rollout=False,
            what="BASELINE.json configs[4]: 256 economies of 10x size, heuristic firms"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--economies", type=int, default=0, help="override the config's economy count")
    ap.add_argument("--burnin", type=int, default=BURNIN)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-staged", action="store_true",
                    help="e2e arm with explicit H2D / D2H copies through the staging block instead of the kernel "
                         "working on the pinned host buffers in place (A/B aid)")
    return ap.parse_args()


def source_hash() -> str:
    """Hash of the sources (and build flags) of the kernel bench.py times at config 3 -- the default-configuration
    one-warp kernel -- to tie profiles/roofline_traffic.json (an ncu capture) to the build that is running."""
    h = hashlib.sha1()
    csrc = ROOT / "r_mabm_b200" / "csrc"
    for name in ("cats_period.cuh", "cats_device.cuh", "cats_layout.h", "cats_inst_static.cu", "build.sh"):
        h.update(name.encode()); h.update((csrc / name).read_bytes())
    return h.hexdigest()[:16]


def workload_config(args, cfg, world, per_gpu, total, blob_bytes, flush):
    state_mb = per_gpu * blob_bytes / 1e6
    return {
        "workload": f"config {args.config} -- {cfg['what']}; {total} economies in all ({per_gpu} on this GPU), "
                    f"W={cfg['W']},F={cfg['F']},N={cfg['N']}, default ABCredit parameters, {args.burnin}-period burn-in "
                    f"untimed, 1 period per step",
        "config": args.config, "economies_total": total, "economies_per_gpu": per_gpu, "periods_per_step": 1,
        "W": cfg["W"], "F": cfg["F"], "N": cfg["N"],
        "parallelism": f"economy-sharded x{world} ({args.scaling} scaling), no step-path collective",
        "l2_policy": (f"state per GPU ({state_mb:.0f} MB) is below the 126 MB L2: a 256 MB buffer is written between "
                      f"timed steps (outside the per-step events)") if flush else
                     f"state per GPU ({state_mb:.0f} MB) exceeds L2 (126 MB); no flush",
    }


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self._nv = None
        try:   # NVML set up here, outside the timed region: the thread only polls
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(index)
            bits = [nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap]
            self._nv = (nv, h, bits, nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _run_nvml(self) -> bool:
        if self._nv is None:
            return False
        nv, h, bits, mx = self._nv
        while not self._halt.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b in bits])
            except Exception:
                pass
            self._halt.wait(0.01)
        return True

    def run(self):
        if self._run_nvml():
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_sample(cfg, burnin, threads, target_s=12.0):
    """The CPU oracle (a port of the model spec -- the reference's Julia path cannot run here) on the box's host
    cores, a bounded sample of the config's workload.  Returns (all-thread rate, one-thread rate, description)."""
    from oracle import oracle as orc
    W, F, N = cfg["W"], cfg["F"], cfg["N"]
    scale = (W + F + N) / 1120.0                         # an economy-period costs about this many default-size ones
    n_econ = max(threads, min(256, 8 * threads))
    burn = burnin if scale < 2 else min(burnin, 60)
    econs = [orc.OracleEconomy(W, F, N, seed=12345, economy=i) for i in range(n_econ)]
    orc.run_batch(econs, burn, n_threads=threads, want_stats=False)
    est = 6.0e3 / scale * threads * 0.4                  # rough rate, to size the sample
    periods = max(5, min(400, int(target_s * est / n_econ)))
    t0 = time.perf_counter()
    orc.run_batch(econs, periods, n_threads=threads, want_stats=False)
    dt = time.perf_counter() - t0
    p1 = max(3, min(50, int(2.0 * 6.0e3 / scale / 4)))
    t0 = time.perf_counter()
    orc.run_batch(econs[:4], p1, n_threads=1, want_stats=False)
    one = 4 * p1 / (time.perf_counter() - t0)
    desc = (f"{n_econ} economies (W={W},F={F},N={N}) x {periods} periods after {burn} burn-in periods, C oracle port of the "
            f"model spec (gcc -O2), OpenMP over economies, {threads} threads; single_thread_value: 4 of them x {p1} "
            f"periods on 1 thread (the reference steps ONE economy on ONE thread through Python + juliacall)")
    return n_econ * periods / dt, one, desc


def cpu_baseline(cfg, burnin=BURNIN, threads=None):
    threads = threads or os.cpu_count() or 1
    rate, one, desc = cpu_sample(cfg, burnin, threads)
    return {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "single_thread_value": one, "sample": desc}


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's own CPU implementation cannot execute in this image (Julia
    + ABCredit.jl absent, no network), so this arm times the CPU oracle port with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    cfg = CONFIGS[args.config]
    W, F, N = cfg["W"], cfg["F"], cfg["N"]
    threads = os.cpu_count() or 1
    scale = (W + F + N) / 1120.0
    n_econ = 8 * threads if scale < 2 else 2 * threads
    periods = 10 if scale < 2 else 3
    burn = args.burnin if scale < 2 else min(args.burnin, 60)
    econs = [orc.OracleEconomy(W, F, N, seed=12345, economy=i) for i in range(n_econ)]
    orc.run_batch(econs, burn, n_threads=threads, want_stats=False)
    for _ in range(args.warmup):
        orc.run_batch(econs, periods, n_threads=threads, want_stats=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.run_batch(econs, periods, n_threads=threads, want_stats=False)
    dt = time.perf_counter() - t0
    value = n_econ * periods * args.steps / dt
    sample = (f"each step = {n_econ} economies (W={W},F={F},N={N}) x {periods} periods (after {burn} burn-in), C oracle "
              f"port of the model spec, OpenMP over economies, {threads} threads")
    total = args.economies or cfg["economies"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"config {args.config} -- {cfg['what']} -- CPU arm: a bounded sample of it, {sample}. "
                                   f"Economies are independent, so the rate per economy-period does not depend on how many "
                                   f"of the {total} economies are stepped",
                       "config": args.config, "economies_total": total, "economies_stepped_per_step": n_econ,
                       "periods_per_step": periods, "W": W, "F": F, "N": N,
                       "parallelism": f"OpenMP over economies, {threads} host threads"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference's juliacall/ABCredit.jl path cannot run in this image; CPU oracle port timed instead"}
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from r_mabm_b200 import _native as nat
    from r_mabm_b200.rollout import shard_range
    from r_mabm_b200.sharded import ShardedCats

    cfg = CONFIGS[args.config]
    W, F, N = cfg["W"], cfg["F"], cfg["N"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on fd 1 at the first collective: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    base = args.economies or cfg["economies"]
    total = base if args.scaling == "strong" else base * world
    nA = cfg["n_agents"]
    sh = ShardedCats(total, rank=rank, world=world, device=dev, W=W, F=F, N=N, n_agents=nA, seed=20260119,
                     log_mode=cfg["rollout"])
    env = sh.env
    B = env.B
    assert (sh.begin, sh.end) == shard_range(total, rank, world)
    lib = nat.load()
    flush = B * env.blob_bytes < 1.2 * L2_BYTES
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush else None

    def flush_l2():
        if flush_buf is not None:
            flush_buf.zero_()

    rollout = None
    if cfg["rollout"]:
        from r_mabm_b200.environments import CatsLog
        from r_mabm_b200.rollout import BatchedRollout, FusedQPolicy
        policy = FusedQPolicy(B, nA, dict(CatsLog._default_gym_spaces_bounds), epsilon=0.3, device=dev, seed=7,
                              economy_base=sh.begin)
        rollout = BatchedRollout(env, policy, t_burnin=args.burnin, T=10 ** 9)
        rollout.reset()

        def step_dev():
            rollout.step(train=True)
    else:
        env.run(args.burnin)                               # the reference's reset(): t_burnin periods
        actions = torch.ones((B, nA, 2), dtype=torch.float64, device=dev)
        mask = torch.zeros((B, nA), dtype=torch.uint8, device=dev)     # "dummy" agents: heuristic decision kept
        h_actions, h_mask = env.host_action_buffers()
        h_actions.fill_(1.0); h_mask.zero_()

        def step_dev():
            env.step(actions, mask, want_stats=True)

    # ---- device-resident arm: CUDA events around every step ----
    W_ = max(args.warmup, 3)
    for _ in range(W_):
        flush_l2(); step_dev()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    launches0 = int(lib.cats_launch_count())
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    t_region = time.perf_counter()
    for i in range(args.steps):
        flush_l2()
        starts[i].record()
        step_dev()
        stops[i].record()
    barrier()
    t_region = time.perf_counter() - t_region
    launches = int(lib.cats_launch_count()) - launches0
    per_launch_ms = [starts[i].elapsed_time(stops[i]) for i in range(args.steps)]
    total_ms = sum(per_launch_ms)
    clocks = sampler.stop() if sampler else None

    # ---- end-to-end arm: host buffers through the public API ----
    if rollout is not None:
        def step_e2e():
            r, _ = rollout.step(train=True)
            return float(r.mean())                          # the step's result read on the host (8 bytes)
        h2d, d2h = 0, 8
    else:
        def step_e2e():
            return env.step_host(h_actions, h_mask, want_stats=True)
        h2d, d2h = env.h2d_bytes_per_step() + B * nA, env.d2h_bytes_per_step(True)
    if args.e2e_staged:
        lib.cats_debug_set_host_zero_copy(int(os.environ.get("CATS_ZC_MODE", "0")))
    for _ in range(3):
        flush_l2(); step_e2e()
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush_l2()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        step_e2e()
        torch.cuda.synchronize(dev)
        e2e_s += time.perf_counter() - t0

    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])

    # the only NCCL traffic of the path: gather the logged statistics once, off the timed region
    last_stats = env.run(1, want_stats=True)
    all_stats = sh.gather(last_stats)
    un_mean = float(all_stats[:, 0, nat.STAT_NAMES.index("Un")].mean())

    # The other scaling mode beside it (N > 1, heuristic configs): the driver's line is the BASELINE-named strong
    # split; the same command also measures `economies` PER GPU (weak), shorter, so that both sit in one record.
    also = None
    if world > 1 and not cfg["rollout"]:
        other = "weak" if args.scaling == "strong" else "strong"
        total2 = base * world if other == "weak" else base
        flush_buf = None                                     # give the 256 MB back
        sh2 = ShardedCats(total2, rank=rank, world=world, device=dev, W=W, F=F, N=N, n_agents=nA, seed=20260119)
        env2, B2 = sh2.env, sh2.env.B
        env2.run(args.burnin)
        a2 = torch.ones((B2, nA, 2), dtype=torch.float64, device=dev)
        m2 = torch.zeros((B2, nA), dtype=torch.uint8, device=dev)
        fl2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if B2 * env2.blob_bytes < 1.2 * L2_BYTES else None
        n2 = max(20, min(args.steps, 100))
        for _ in range(5):
            env2.step(a2, m2, want_stats=True)
        barrier()
        s2 = [torch.cuda.Event(enable_timing=True) for _ in range(n2)]
        e2 = [torch.cuda.Event(enable_timing=True) for _ in range(n2)]
        for i in range(n2):
            if fl2 is not None:
                fl2.zero_()
            s2[i].record(); env2.step(a2, m2, want_stats=True); e2[i].record()
        barrier()
        t2 = torch.tensor([sum(s2[i].elapsed_time(e2[i]) for i in range(n2))], dtype=torch.float64, device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        also = {"scaling": other, "economies_total": total2, "economies_per_gpu": B2, "steps": n2,
                "value": total2 * n2 / (float(t2[0]) * 1e-3), "unit": UNIT,
                "l2_flush_between_steps": fl2 is not None}

    if rank == 0:
        value = total * args.steps / (total_ms * 1e-3)
        peak, peak_src = measured_peak_hbm()
        algo = env.algo_bytes_per_period() * B              # bytes per launch on this GPU
        avg_launch_s = statistics.fmean(per_launch_ms) * 1e-3
        achieved = algo / avg_launch_s / 1e9
        traffic, issue, stale = None, None, None
        tp = ROOT / "profiles" / "roofline_traffic.json"
        if tp.exists() and args.config == 3:
            try:
                tj = json.loads(tp.read_text())
                stale = tj.get("source_hash") != source_hash()
                if not stale:
                    traffic = tj.get("dram_bytes_per_economy_period", 0) * B or None
                    wi = tj.get("warp_instructions_per_economy_period")
                    if wi and clocks and clocks.get("sm_mhz"):
                        sms = env.launch_info()["num_sms"]
                        peak_issue = sms * 4 * clocks["sm_mhz"] * 1e6
                        issue = {"warp_instructions_per_economy_period": wi, "achieved_ginst_s": B / avg_launch_s * wi / 1e9,
                                 "peak_ginst_s": peak_issue / 1e9, "frac": (B / avg_launch_s * wi) / peak_issue,
                                 "source": "profiles/roofline_traffic.json (ncu, same kernel sources) x rate measured in this run"}
            except Exception:
                traffic = None
        srt = sorted(per_launch_ms)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": W_, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, cfg, world, B, total, env.blob_bytes, flush),
            "e2e": {"value": total * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "transport": ("device-resident policy; the step's mean reward read on the host" if rollout is not None
                                  else "explicit cudaMemcpyAsync through a device staging block" if args.e2e_staged
                                  else "actions: one pinned H2D copy; outputs: written in place into pinned host buffers by the kernel over PCIe (cats_step_host)")},
            "gpu_launches": launches,
            "step_ms": {"min": srt[0], "median": srt[len(srt) // 2], "max": srt[-1], "mean": statistics.fmean(srt)},
            "timed_region_s": t_region,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_stale": stale, "peak_source": peak_src,
                         "algo_bytes_per_economy_period": env.algo_bytes_per_period(),
                         "issue_slots": issue,
                         "note": "bound by dependent-instruction latency and issue slots of the serial matching "
                                 "markets, not by bandwidth; see DESIGN.md 2.3"},
            "clocks": clocks, "launch": env.launch_shape(), "mean_unemployment_check": un_mean,
        }
        if also is not None:
            line["also_measured"] = also
        if not args.no_cpu_baseline and world == 1:        # reported at N = 1 only (the host cores are busy driving N ranks otherwise)
            cb = cpu_baseline(cfg, args.burnin)
            line["cpu_baseline"] = cb
            line["vs_cpu_single_thread"] = value / cb["single_thread_value"]
            line["vs_cpu_all_threads"] = value / cb["value"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
