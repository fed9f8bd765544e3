"""Summarise the ncu outputs of tools/ncu_round.sh into profiles/ (tracked).

  python tools/ncu_summary.py <tag> <workload>
reads  gpurun_out/<tag>_launches.csv  (ncu --metrics gpu__time_duration.sum launch list)
       gpurun_out/<tag>_bulk.ncu-rep  (ncu --set full capture of the bulk kernel)
writes profiles/<tag>_launches.csv    (the launch list, verbatim)
       profiles/<tag>_launches_summary.md  (per-kernel count / total / share of the step)
       profiles/<tag>_bulk_raw.csv    (selected raw metrics per captured launch)
       profiles/bulk_traffic.json     (dram bytes per launch, read by bench.py's roofline.traffic)
"""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, workload = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "slab")
suffix = sys.argv[3] if len(sys.argv) > 3 else "bulk"  # gpurun_out/<tag>_<suffix>.ncu-rep
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "lts__t_bytes.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__occupancy_limit_shared_mem", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("unnamed>::", "").replace("lbm::", "")
    return re.sub(r"\(.*$", "", name)


# ---- launch list ------------------------------------------------------------------------------
src = os.path.join(G, tag + "_launches.csv")
if os.path.exists(src) and suffix == "bulk":
    shutil.copy(src, os.path.join(P, tag + "_launches.csv"))
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    with open(os.path.join(P, tag + "_launches_summary.md"), "w") as f:
        f.write("# %s: every kernel launch of `python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity` (workload %s)\n\n" % (tag, workload))
        f.write("ncu `--metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and\n"
                "serialised, so compare SHARES.  %d launches, %.3f ms of kernel time in total.\n\n" % (len(rows), total / 1e6))
        f.write("| kernel | launches | total ms | avg us | share | grid | block |\n|---|---:|---:|---:|---:|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f | %.1f %% | %s | %s |\n" % (k, a[0], a[1] / 1e6, a[1] / a[0] / 1e3, 100 * a[1] / total, a[2], a[3]))
        step = {k: a for k, a in agg.items() if k.startswith(("k_tb", "k_bulk", "k_fixup", "k_forces", "k_wrap"))}
        st = sum(a[1] for a in step.values())
        if st:
            f.write("\nStep kernels only (what the timed region of bench.py launches):\n\n")
            for k, a in sorted(step.items(), key=lambda kv: -kv[1][1]):
                f.write("* `%s`: %.1f %% of step-kernel time (%d launches, avg %.1f us)\n" % (k, 100 * a[1] / st, a[0], a[1] / a[0] / 1e3))
    print(open(os.path.join(P, tag + "_launches_summary.md")).read())

# ---- full capture -----------------------------------------------------------------------------
rep = os.path.join(G, tag + "_" + suffix + ".ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    all_data = data
    cols = [head.index(k) for k in KEEP if k in head]
    with open(os.path.join(P, tag + "_" + suffix + "_raw.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [head[c] for c in cols])
        w.writerow(["unit"] + [units[c] for c in cols])
        for d in all_data:
            w.writerow([short(d[head.index("Kernel Name")])] + [d[c] for c in cols])
    # the traffic figure is that of the dominant (bulk collide-stream) kernel only
    data = [d for d in all_data if re.search(r"k_tb|k_bulk|k_aa_(odd|even)", d[head.index("Kernel Name")])]

    def col(name):
        c = head.index(name)
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(units[c], 1.0)
        return [float(d[c].replace(",", "")) * scale for d in data]

    rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
    per_launch = sum(a + b for a, b in zip(rd, wr)) / len(rd)
    tj = os.path.join(P, "bulk_traffic.json")
    j = json.load(open(tj)) if os.path.exists(tj) else {}
    j[workload] = {"dram_bytes_per_launch": per_launch, "dram_read": sum(rd) / len(rd), "dram_write": sum(wr) / len(wr),
                   "launches_captured": len(rd), "kernel": " / ".join(sorted({short(d[head.index("Kernel Name")]) for d in data})),
                   "source": "profiles/%s_%s_raw.csv (ncu --set full --clock-control none)" % (tag, suffix)}
    json.dump(j, open(tj, "w"), indent=1)
    print(json.dumps(j[workload], indent=1))
    # stall reasons per source line (top 12 lines by samples)
    srcp = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    open(os.path.join(P, tag + "_" + suffix + "_source.csv"), "w").write(srcp)
