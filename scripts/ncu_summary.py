"""Turn ncu output into the text summaries committed under profiles/.

  python scripts/ncu_summary.py raw      <report.ncu-rep> <out.txt> ["header comment"]
  python scripts/ncu_summary.py traffic  <report.ncu-rep> <out.json> <particles> ["source note"]
  python scripts/ncu_summary.py launches <launches.csv>   <out.txt> <steps> ["header comment"]

`raw`: the metrics the roofline discussion of DESIGN.md section 3 quotes, one row per metric,
one column per captured launch.  `traffic`: dram__bytes_read.sum + dram__bytes_write.sum per
launch of the density and force sweeps (what bench.py reports as roofline.traffic).
`launches`: per-kernel share of the step from a `--metrics gpu__time_duration.sum` pass.
"""

import csv
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size",
    "sm__cycles_elapsed.max", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def raw_rows(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], check=True,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def short(name):
    return name.replace("sphb200::", "").replace("(int)", "")[:44]


def cmd_raw(report, dst, note=""):
    hdr, units, data = raw_rows(report)
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {note}"] if note else []
    lines.append(f"Kernel Name  {[short(d[col['Kernel Name']]) for d in data]}")
    for m in METRICS:
        if m in col:
            lines.append(f"{m} {units[col[m]]} {[d[col[m]] for d in data]}")
    open(dst, "w").write("\n".join(lines) + "\n")


def cmd_traffic(report, dst, n, note=""):
    """n: particles of the captured run; the file also carries bench.source_sha(), the hash of
    the CUDA sources, so that bench.py only quotes it for the build it describes."""
    import os

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    n = int(n)
    hdr, units, data = raw_rows(report)
    col = {h: i for i, h in enumerate(hdr)}

    def to_bytes(v, unit):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
        return float(v.replace(",", "")) * scale

    out = {}
    for d in data:
        name = d[col["Kernel Name"]]
        key = "density" if "PhysDensity" in name else "force" if "PhysForce" in name else None
        if key is None or key in out:
            continue
        rd = to_bytes(d[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(d[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        out[key] = {"kernel": name.replace("sphb200::", "").replace("(int)", ""),
                    "dram_bytes_read": rd, "dram_bytes_write": wr,
                    "dram_bytes_per_launch": rd + wr, "particles": n,
                    "dram_bytes_per_particle": (rd + wr) / n}
    out["_source"] = note
    out["source_sha"] = bench.source_sha()
    json.dump(out, open(dst, "w"), indent=1)


def cmd_launches(src, dst, steps, note=""):
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
    per = {}
    for r in rows:
        k = short(r[4].split("(")[0] if not r[4].startswith("void") else r[4][5:].split("(")[0])
        per[k] = per.get(k, 0.0) + float(r[14].replace(",", "")) / 1e6
    tot = sum(per.values())
    lines = [f"# {note}"] if note else []
    lines.append(f"# {len(rows)} launches = {steps} timed steps; cold-cache serialised times: compare SHARES")
    lines.append(f"# total {tot:.3f} ms for {steps} steps")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1]):
        lines.append(f"{100 * v / tot:6.2f}%  {v / steps:10.3f} ms/step  {k}")
    open(dst, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "raw":
        cmd_raw(*sys.argv[2:5])
    elif mode == "traffic":
        cmd_traffic(*sys.argv[2:6])
    elif mode == "launches":
        cmd_launches(sys.argv[2], sys.argv[3], int(sys.argv[4]), *sys.argv[5:6])
