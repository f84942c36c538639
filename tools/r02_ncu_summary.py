"""Writes profiles/r02_ncu_summary.md from the round-2 captures under gpurun_out/: the launch list of the bench command
(r02_launches_c3.csv.gz, `--metrics gpu__time_duration.sum --clock-control none`) and the `--set full` captures of the dominant
kernels (r02_prof_tc_xt / r02_prof_syrk / r02_prof_chain .ncu-rep).  Usage: python tools/r02_ncu_summary.py"""
import csv, gzip, io, os, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
out = ["# profiles/r02 — ncu evidence, round 2 (B200, gpurun fresh box, `--clock-control none`)", "",
       "Captured by `tools/r02_capture_final.sh` on the final code of the round.  Per-launch times under ncu are cold-cache and "
       "serialised: compare SHARES, not absolutes; numbers quoted as measured come from `profiles/r02_bench_c3.json` (plain run).", ""]
p = os.path.join(G, "r02_launches_c3.csv.gz")
if os.path.exists(p):
    rows = list(csv.reader(io.TextIOWrapper(gzip.open(p), errors="replace")))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    col = {c: i for i, c in enumerate(rows[h])}
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[col["Metric Value"]].replace(",", "")); u = r[col["Metric Unit"]]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
        k = r[col["Kernel Name"]]; agg[k][0] += 1; agg[k][1] += us
    tot = sum(v[1] for v in agg.values())
    out += ["## 1. launch list — `r02_launches_c3.csv.gz`  (`COVINS_SKIP_CPU_BASELINE=1 ncu --metrics gpu__time_duration.sum --clock-control none … "
            "python bench.py --steps 2 --warmup 3`: GBA at C3 incl. the e2e call, PGO at C2, ORB / SIFT matching, microbenchmarks)", "",
            f"total {tot/1e3:.1f} ms over {sum(v[0] for v in agg.values())} launches", "", "| kernel | launches | sum ms | share | avg us |", "|---|---|---|---|---|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        out.append(f"| `{k[:100]}` | {n} | {us/1e3:.2f} | {100*us/tot:.1f} % | {us/n:.1f} |")
    out.append("")
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
           "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
sec = 2
for rep, title in (("r02_prof_tc_xt", "`cvb_tc::xt::tc_xt_kernel<2>` (K1; C3 request: 1000 queries x 2000 KF x 1000 rows = 2 Gpairs per launch; `tools/tc_profile.py`)"),
                   ("r02_prof_chain", "`cvb_chol::chain_gemm_kernel` (K8 critical chain: solve of the first panel tile / update of the next diagonal tile)")):
    f = os.path.join(G, rep + ".ncu-rep")
    if not os.path.exists(f):
        continue
    txt = subprocess.check_output(["ncu", "-i", f, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out += [f"## {sec}. `ncu --set full` — {title}", ""]; sec += 1
    for r in rows[2:4]:
        out += ["| metric | value | unit |", "|---|---|---|", f"| Kernel Name | {r[col['Kernel Name']][:110]} |  |"]
        for m in METRICS:
            if m in col and r[col[m]] != "":
                out.append(f"| {m} | {r[col[m]]} | {units[col[m]]} |")
        out.append("")
open(os.path.join(ROOT, "profiles", "r02_ncu_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:60]))
