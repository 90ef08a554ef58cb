#!/usr/bin/env python3
"""Digest an .ncu-rep (read here, no GPU needed): headline counters + per-source-line instruction and stall shares.

    python profiles/ncu_digest.py gpurun_out/prof.ncu-rep [--units N] [--top K] [--json out.json]

--units N divides the instruction count by N (e.g. check-node updates in the profiled launch).
"""
import argparse
import csv
import io
import json
import subprocess
import sys

HEAD = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_adu.sum", "sm__inst_executed_pipe_cbu.sum"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--units", type=float, default=0)
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--json")
    a = ap.parse_args()
    rows = list(csv.reader(io.StringIO(ncu(a.rep, "--page", "raw", "--csv"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {"kernel": vals[hdr.index("Kernel Name")]}
    for h, u, v in zip(hdr, units, vals):
        if h in HEAD or h.startswith("smsp__average_warps_issue_stalled"):
            out[h] = [v, u]
    for h in HEAD:
        if h in out:
            print("%-70s %s %s" % (h, out[h][0], out[h][1]))
    st = sorted(((float(v[0]), h) for h, v in out.items() if h.startswith("smsp__average_warps_issue_stalled")), reverse=True)
    print("stall reasons (warps per issue):", ", ".join("%s=%.2f" % (h.split("stalled_")[1].split("_per")[0], x) for x, h in st[:8]))
    inst = float(out["smsp__inst_executed.sum"][0])
    if a.units:
        out["warp_inst_per_unit"] = inst / a.units
        print("warp instructions per unit: %.1f" % (inst / a.units))
    src = list(csv.reader(io.StringIO(ncu(a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    cur, hd, agg = None, None, {}
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 2 and r[0] == "Line No":
            hd = r
        elif hd and len(r) > 10 and r[0] != "":
            try:
                agg[(cur, int(r[0]))] = (int(r[hd.index("Instructions Executed")]), int(r[hd.index("# Samples")]), r[1].strip()[:100])
            except ValueError:
                pass
    tot = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    lines = []
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
        lines.append({"file": k[0], "line": k[1], "inst_pct": 100 * v[0] / tot, "stall_sample_pct": 100 * v[1] / ts, "src": v[2]})
        print("%-18s %5d  inst %5.1f%%  samples %5.1f%%  %s" % (k[0], k[1], 100 * v[0] / tot, 100 * v[1] / ts, v[2]))
    out["top_lines_by_samples"] = lines
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    sys.exit(main())
