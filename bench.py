#!/usr/bin/env python
"""bench.py -- economy-periods/sec of the batched CATS step on N B200s (one process per GPU).

  python bench.py --gpus 1 --steps 50 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference --gpus 1 --steps 5 --warmup 1     # CPU arm (oracle port)

Workload (BASELINE.json configs[2], the one the metric is quoted on at 1/2/4/8 GPUs): 16,384
default-size economies (W=1000, F=100, N=20, default parameters) PER GPU, Monte-Carlo seeds,
heuristic firms, after the reference's 300-period burn-in.  A *step* is one period of every
economy: one launch of the period kernel through the C-ABI, state round-tripping through HBM
(620 MB of economy state per GPU >> 126 MB of L2, so no L2 flush is needed between steps).
Economies are independent: ranks own disjoint slabs of global economy ids, there is no collective
on the step path (scaling = weak); NCCL is used once, after the timed region, to gather statistics.

One JSON line on stdout (rank 0).  `value` = device-resident inputs, `e2e` = the same step through
the public API with HOST buffers (pinned action/mask upload, obs/reward/terminated/status/stats
download inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

W, F, N = 1000, 100, 20
ECONOMIES_PER_GPU = 16384
BURNIN = 300
METRIC = "economy-periods/sec (batched CATS step)"
UNIT = "economy-periods/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--economies", type=int, default=ECONOMIES_PER_GPU, help="economies per GPU")
    ap.add_argument("--burnin", type=int, default=BURNIN)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(args, world, blob_bytes=37760):
    return {
        "workload": f"{args.economies} default-size economies per GPU (W={W},F={F},N={N}, default ABCredit "
                    f"parameters), Monte-Carlo seeds, heuristic firms (n_agents=1, dummy action mask), "
                    f"{args.burnin}-period burn-in untimed, 1 period per step",
        "economies_per_gpu": args.economies, "economies_total": args.economies * world,
        "periods_per_step": 1, "W": W, "F": F, "N": N,
        "parallelism": f"economy-sharded x{world}, no step-path collective",
        "l2_policy": "state per GPU (%.0f MB) exceeds L2 (126 MB); no flush" % (args.economies * blob_bytes / 1e6),
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
        """Same fields through NVML (a query takes microseconds, so short timed regions get many samples)."""
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


def cpu_baseline(sample_economies=256, sample_periods=200, burnin=BURNIN, threads=None):
    """The CPU oracle (a port of the model spec -- the reference's Julia path cannot run here) on the
    box's host cores, bounded sample of the same workload."""
    from oracle import oracle as orc
    threads = threads or os.cpu_count() or 1
    econs = [orc.OracleEconomy(W, F, N, seed=12345, economy=i) for i in range(sample_economies)]
    orc.run_batch(econs, burnin, n_threads=threads, want_stats=False)
    t0 = time.perf_counter()
    orc.run_batch(econs, sample_periods, n_threads=threads, want_stats=False)
    dt = time.perf_counter() - t0
    # the reference steps ONE economy on ONE thread (Python + juliacall): the like-for-like figure
    t0 = time.perf_counter()
    orc.run_batch(econs[:4], 50, n_threads=1, want_stats=False)
    one = 4 * 50 / (time.perf_counter() - t0)
    return {"value": sample_economies * sample_periods / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "single_thread_value": one,
            "sample": f"{sample_economies} default-size economies x {sample_periods} periods after {burnin} burn-in "
                      f"periods, C oracle (gcc -O2), OpenMP over economies, {threads} threads; single_thread_value: "
                      f"4 of them x 50 periods on 1 thread"}


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's own CPU implementation cannot execute in this image (Julia
    + ABCredit.jl absent, no network), so this arm times the CPU oracle port with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    threads = os.cpu_count() or 1
    n_econ, periods = 8 * threads, 10
    econs = [orc.OracleEconomy(W, F, N, seed=12345, economy=i) for i in range(n_econ)]
    orc.run_batch(econs, args.burnin, n_threads=threads, want_stats=False)
    for _ in range(args.warmup):
        orc.run_batch(econs, periods, n_threads=threads, want_stats=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.run_batch(econs, periods, n_threads=threads, want_stats=False)
    dt = time.perf_counter() - t0
    value = n_econ * periods * args.steps / dt
    sample = (f"each step = {n_econ} default-size economies x {periods} periods (after {args.burnin} burn-in), "
              f"C oracle port of the model spec, OpenMP over economies, {threads} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference's juliacall/ABCredit.jl path cannot run in this image; CPU oracle port timed instead"}
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from r_mabm_b200 import _native as nat
    from r_mabm_b200.engine import BatchedCats

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

    B = args.economies
    env = BatchedCats(B, W=W, F=F, N=N, n_agents=1, seed=20260119, economy_base=rank * B, device=dev)
    env.run(args.burnin)                                   # the reference's reset(): t_burnin periods
    actions = torch.ones((B, 1, 2), dtype=torch.float64, device=dev)
    mask = torch.zeros((B, 1), dtype=torch.uint8, device=dev)     # "dummy" agent: heuristic decision kept
    # pinned host buffers the caller fills: the engine's staging block (one H2D and one D2H copy per step)
    h_actions, h_mask = env.host_action_buffers()
    h_actions.fill_(1.0); h_mask.zero_()
    lib = nat.load()

    # ---- device-resident arm ----
    for _ in range(max(args.warmup, 3)):
        env.step(actions, mask, want_stats=True)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    launches0 = int(lib.cats_launch_count())
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        env.step(actions, mask, want_stats=True)
        evs[i + 1].record()
    barrier()
    launches = int(lib.cats_launch_count()) - launches0
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if sampler else None

    # ---- end-to-end arm: host buffers through the public API ----
    for _ in range(3):
        env.step_host(h_actions, h_mask, want_stats=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        obs, rew, term, stats = env.step_host(h_actions, h_mask, want_stats=True)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])

    # the only NCCL traffic of the path: gather the logged statistics once, off the timed region
    last_stats = env.step(actions, mask, want_stats=True)[3][:, 0, :].contiguous()
    if world > 1:
        gathered = [torch.empty_like(last_stats) for _ in range(world)]
        dist.all_gather(gathered, last_stats)
        all_stats = torch.cat(gathered, 0)
    else:
        all_stats = last_stats
    un_mean = float(all_stats[:, nat.STAT_NAMES.index("Un")].mean())

    if rank == 0:
        value = B * world * args.steps / (total_ms * 1e-3)
        peak, peak_src = measured_peak_hbm()
        algo = env.algo_bytes_per_period() * B              # bytes per launch on this GPU
        avg_launch_s = (sum(per_launch_ms) / len(per_launch_ms)) * 1e-3
        achieved = algo / avg_launch_s / 1e9
        traffic, issue = None, None
        tp = ROOT / "profiles" / "roofline_traffic.json"
        if tp.exists():
            try:
                tj = json.loads(tp.read_text())
                traffic = tj.get("dram_bytes_per_economy_period", 0) * B or None
                # what actually bounds the kernel: warp-instruction issue.  Instruction
                # count per economy-period from the committed ncu capture, rate measured here.
                wi = tj.get("warp_instructions_per_economy_period")
                if wi and clocks and clocks.get("sm_mhz"):
                    sms = env.launch_info()["num_sms"]
                    peak_issue = sms * 4 * clocks["sm_mhz"] * 1e6
                    issue = {"warp_instructions_per_economy_period": wi, "achieved_ginst_s": B / avg_launch_s * wi / 1e9,
                             "peak_ginst_s": peak_issue / 1e9, "frac": (B / avg_launch_s * wi) / peak_issue,
                             "source": "profiles/roofline_traffic.json (ncu) x rate measured in this run"}
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world, env.blob_bytes),
            "e2e": {"value": B * world * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": env.h2d_bytes_per_step() + B, "d2h_bytes_per_step": env.d2h_bytes_per_step(True)},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algo_bytes_per_economy_period": env.algo_bytes_per_period(),
                         "issue_slots": issue,
                         "note": "bound by dependent-instruction latency and issue slots of the serial matching "
                                 "markets, not by bandwidth; see DESIGN.md 2.3"},
            "clocks": clocks, "launch": env.launch_info(), "mean_unemployment_check": un_mean,
        }
        if not args.no_cpu_baseline and world == 1:        # reported at N = 1 only (the host cores are busy driving N ranks otherwise)
            line["cpu_baseline"] = cpu_baseline()
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
