"""Summarises ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.
usage: python tools/summarize_ncu.py <launches.csv> <full.ncu-rep> <tag>"""
import csv
import collections
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def launches(path, tag):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "").replace("b200zk::", "")
        ns = float(r[-1].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    out = ["# ncu launch list summary (%s)" % tag, "",
           "command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 1 "
           "--no-prove --no-sizes --no-cpu-baseline` (input generation, 3 warm-up + 2 timed + 2 end-to-end steps, the pipelined and "
           "fixed-base sections)",
           "(cold-cache, serialised launches: compare SHARES, not absolutes)", "",
           "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.3f | %.1f%% |" % (k, n, ns / 1e6, 100 * ns / total))
    open(os.path.join(ROOT, "profiles", "%s_launches.md" % tag), "w").write("\n".join(out) + "\n")


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def full(path, tag):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ["# ncu --set full summary (%s): %s" % (tag, os.path.basename(path)), ""]
    traffic = {}
    for r in rows[2:]:
        kname = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        out.append("## launch %s  `%s`" % (r[0], kname.replace("b200zk::", "")[:90]))
        vals = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals[k] = (r[i], units[i])
                out.append("- %s = %s %s" % (k, r[i], units[i]))
        try:
            def mb(x):
                v, u = vals[x]
                f = float(v.replace(",", ""))
                return f * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u]
            traffic[kname] = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
            out.append("- **DRAM traffic per launch = %.1f MB**" % (traffic[kname] / 1e6))
        except Exception:
            pass
        out.append("")
    open(os.path.join(ROOT, "profiles", "%s_full.md" % tag), "w").write("\n".join(out) + "\n")
    return traffic


if __name__ == "__main__":
    lpath, fpath, tag = sys.argv[1:4]
    if os.path.exists(lpath):
        launches(lpath, tag)
    if os.path.exists(fpath):
        t = full(fpath, tag)
        acc = [v for k, v in t.items() if "k_msm_accumulate" in k]
        if acc:
            json.dump({"msm_accumulate_g1_bytes_per_launch": sum(acc) / len(acc), "source": tag},
                      open(os.path.join(ROOT, "profiles", "traffic.json"), "w"))
