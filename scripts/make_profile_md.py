"""profiles/<round>_ncu_<name>.md (+ roofline_traffic.json for the bench kernel) from an ncu --set full capture.
usage: make_profile_md.py <rep> <so> <kernel-substring> <units in the launch> <round tag> [name [what [unit]]]
  name: file suffix (default period_kernel: also writes profiles/roofline_traffic.json, tied to the kernel sources by hash)
  what: one line describing the launch;  unit: what a "unit" is (default economy-period)"""
import csv, json, re, subprocess, sys
sys.path.insert(0, ".")
rep, so, kern, nep, tag = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), sys.argv[5]
name = sys.argv[6] if len(sys.argv) > 6 else "period_kernel"
what = sys.argv[7] if len(sys.argv) > 7 else f"{int(nep)} default-size economies, ONE period, after the 300-period burn-in."
unit = sys.argv[8] if len(sys.argv) > 8 else "economy-period"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, r = rows[0], rows[1], rows[2]
if kern == "auto":
    # the exact mangled name of the kernel that was captured, from its demangled name in the report (a substring such
    # as "ILi5ELi14ELi2ELb0ELi2" also matches the strict-mode instantiation and mixes two kernels' line tables)
    dem = r[h.index("Kernel Name")]
    mt = re.search(r"(cats_period_mw_kernel|cats_period_kernel)<([^>]*)>", dem)
    if mt:
        kinds = "iiibib" if mt.group(1) == "cats_period_kernel" else "iiiib"
        args = [a.strip() for a in mt.group(2).split(",")]
        kern = mt.group(1) + "I" + "".join(("Lb" if k == "b" else "Li") + a + "E" for k, a in zip(kinds, args)) + "E"
    else:
        kern = re.search(r"(cats_\w+)", dem).group(1)
    print("kernel:", dem, "->", kern)
m = {n: (r[i], u[i]) for i, n in enumerate(h)}
def f(name): return float(m[name][0].replace(",", ""))
def unit_bytes(name):
    v, un = f(name), m[name][1]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[un]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
dur_us = f("gpu__time_duration.sum") * {"us": 1, "ms": 1e3, "ns": 1e-3, "second": 1e6}[m["gpu__time_duration.sum"][1]]
dr, dw = unit_bytes("dram__bytes_read.sum"), unit_bytes("dram__bytes_write.sum")
inst = f("smsp__inst_executed.sum")
out = [f"# ncu --set full: {name} ({kern}), {tag}", "",
       "Capture: scripts/r3_profile.sh (gpurun, 1 B200, each command directly after the same command exited 0 without ncu;",
       "`ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1`).", "",
       f"Launch profiled: {what}",
       "Times under ncu are cold-cache and serialised: use shares, not absolutes.", "", "| metric | value |", "|---|---|"]
for n in want:
    if n in m: out.append(f"| {n} | {m[n][0]} {m[n][1]} |")
out += [f"| warp instructions per {unit} | {inst / nep:.0f} |",
        f"| DRAM bytes per {unit} (read + write) | {(dr + dw) / nep:.0f} |",
        f"| {unit}s/s under ncu | {nep / (dur_us * 1e-6):.3e} |", f"| achieved DRAM GB/s under ncu (read + write) | {(dr + dw) / (dur_us * 1e-6) / 1e9:.0f} |", "",
        "Stall reasons (cycles per issued instruction, per warp):", "", "| reason | cycles |", "|---|---|"]
st = [(float(r[i]), n) for i, n in enumerate(h) if re.fullmatch(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", n)]
for v, n in sorted(st, reverse=True):
    if v >= 0.005: out.append(f"| {n.split('stalled_')[1].split('_per_issue')[0]} | {v:.2f} |")
by = subprocess.run([sys.executable, "scripts/ncu_stalls_by_region.py", rep, so, kern, str(nep)], capture_output=True, text=True).stdout
out += ["", f"Per source function (instructions per {unit}, share of stall samples, stall mix in % of all samples):", "", "```", by.rstrip(), "```"]
open(f"profiles/{tag}_ncu_{name}.md", "w").write("\n".join(out) + "\n")
if name != "period_kernel":
    print("\n".join(out[:40])); sys.exit(0)
from bench import source_hash
json.dump({"kernel": kern, "source": f"profiles/{tag}_ncu_period_kernel.md (ncu --set full, {int(nep)} economies x 1 period)",
           "source_hash": source_hash(),
           "dram_bytes_read": dr, "dram_bytes_write": dw, "economy_periods": nep,
           "dram_bytes_per_economy_period": (dr + dw) / nep,
           "warp_instructions_per_economy_period": inst / nep,
           "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active")},
          open("profiles/roofline_traffic.json", "w"), indent=1)
print("\n".join(out[:48]))
