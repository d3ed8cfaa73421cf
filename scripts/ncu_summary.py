#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small JSON for profiles/: per profiled launch the duration,
registers, occupancy, DRAM bytes, pipe utilisations, issue activity, stall ratios and the executed SASS opcode mix.

    python scripts/ncu_summary.py gpurun_out/r01_prof.ncu-rep profiles/r01_ncu_eval_tape.json
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def ncu(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True, check=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in body:
        d = {"kernel": r[idx["Kernel Name"]]}
        for k in KEYS:
            if k in idx:
                try:
                    d[k] = float(r[idx[k]].replace(",", ""))
                except ValueError:
                    d[k] = r[idx[k]]
                d.setdefault("_units", {})[k] = units[idx[k]]
        launches.append(d)
    # executed opcode mix per launch from the source page
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    heads = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    for li, hi in enumerate(heads):
        h = src[hi]
        end = heads[li + 1] - 1 if li + 1 < len(heads) else len(src)
        ci, si = h.index("Instructions Executed"), h.index("Source")
        mix = collections.Counter()
        for r in src[hi + 1:end]:
            if len(r) <= ci:
                continue
            t = r[si].split()
            if not t:
                continue
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            op = op.rstrip(";")
            if op.startswith("IMAD.WIDE"):
                key = "IMAD.WIDE"
            elif op.startswith("IMAD.MOV"):
                key = "IMAD.MOV"
            elif op.startswith("IMAD.HI"):
                key = "IMAD.HI"
            elif op.startswith("IMAD"):
                key = "IMAD"
            else:
                key = op.split(".")[0]
            mix[key] += int(r[ci])
        tot = sum(mix.values())
        if li < len(launches):
            launches[li]["warp_instructions"] = tot
            launches[li]["opcode_mix_pct"] = {k: round(100.0 * v / tot, 2) for k, v in mix.most_common(14)}
    with open(out, "w") as f:
        json.dump({"report": rep, "launches": launches}, f, indent=1)
    for d in launches:
        print(d["kernel"][:60], d.get("gpu__time_duration.sum"), d.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"))


if __name__ == "__main__":
    main()
