"""Summarise ncu reports (gpurun_out/*.ncu-rep) into tracked text/JSON files under profiles/.

    python profiles/summarize.py <tag> <rep1.ncu-rep> [<rep2.ncu-rep> ...]

Writes profiles/<tag>_summary.json (+ .txt) and refreshes profiles/latest_traffic.json, the
per-launch DRAM traffic bench.py reports in roofline.traffic.
"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")]}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            try:
                v = float(vals[i].replace(",", ""))
            except ValueError:
                continue
            if units[i] in UNIT:
                v *= UNIT[units[i]]
                d[w + " [B]"] = v
            else:
                d[w + (f" [{units[i]}]" if units[i] else "")] = v
    return d


def hot_lines(rep, top=14):
    """per CUDA source line: warp instructions executed and stall samples, summed over the SASS instructions
    ncu lists under that line (the cuda,sass view carries the counters on the SASS rows only)"""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    if not hi:
        return []
    h = rows[hi[0]]
    cs, ci = h.index("# Samples"), h.index("Instructions Executed")
    agg, cur = {}, None
    for r in rows[hi[0] + 1:]:
        if len(r) <= max(cs, ci):
            continue
        if r[0] != "":
            cur = (r[0], r[1].strip()[:100])
            agg.setdefault(cur, [0.0, 0.0])
            continue
        if cur is None or r[2] in ("", "..."):
            continue
        try:
            agg[cur][0] += float(r[cs]); agg[cur][1] += float(r[ci])
        except ValueError:
            continue
    tot = sum(v[0] for v in agg.values()) or 1.0
    tin = sum(v[1] for v in agg.values()) or 1.0
    best = sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]
    return [{"pct_samples": round(100 * v[0] / tot, 1), "pct_inst": round(100 * v[1] / tin, 1), "warp_inst": int(v[1]),
             "line": int(k[0]) if k[0].isdigit() else 0, "src": k[1]} for k, v in best]


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    here = os.path.dirname(os.path.abspath(__file__))
    summary, traffic = [], {}
    for rep in reps:
        d = raw(rep)
        d["report"] = os.path.basename(rep)
        d["hot_source_lines"] = hot_lines(rep)
        summary.append(d)
        t = d.get("dram__bytes_read.sum [B]", 0.0) + d.get("dram__bytes_write.sum [B]", 0.0)
        for key in ("hermite", "leaves", "rows", "columns", "solve"):
            if "k_" + key in d["kernel"]:
                traffic[key + "_dram_bytes_per_launch"] = t
                # executed warp instructions of that launch: bench.py's issue-slot fraction
                traffic[key + "_warp_inst_per_launch"] = d.get("smsp__inst_executed.sum [inst]", 0.0)
    json.dump(summary, open(os.path.join(here, f"{tag}_summary.json"), "w"), indent=1)
    with open(os.path.join(here, f"{tag}_summary.txt"), "w") as f:
        for d in summary:
            f.write(f"== {d['kernel'][:90]}  ({d['report']})\n")
            for k, v in d.items():
                if k not in ("kernel", "report", "hot_source_lines"):
                    f.write(f"   {k:72s} {v:,.3f}\n")
            for hl in d["hot_source_lines"]:
                f.write(f"   {hl['pct_inst']:5.1f}% inst {hl['pct_samples']:5.1f}% samples  L{hl['line']}: {hl['src']}\n")
            f.write("\n")
    traffic["source"] = f"profiles/{tag}_summary.json (ncu --set full --clock-control none, one launch each)"
    json.dump(traffic, open(os.path.join(here, "latest_traffic.json"), "w"), indent=1)
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
