#!/usr/bin/env python
"""Turn an ncu report (.ncu-rep) into the short summary committed under profiles/.

usage: python profiles/summarize_ncu.py gpurun_out/X.ncu-rep "<title>" > profiles/X.md
Reads `ncu -i ... --page raw --csv` and `--page source --csv` (no GPU needed)."""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("sm__warps_active.avg.per_cycle_active", "warps active per SM (avg)"),
    ("smsp__issue_active.avg.per_cycle_active", "issue slots used per SMSP cycle"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe utilisation (active cycles)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "FP64 pipe utilisation (elapsed)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per instruction"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (unused by design)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: fixed-latency dependency (wait)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, title = sys.argv[1], sys.argv[2]
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    print(f"source: `{rep}` (ncu --set full --clock-control none), kernel `{d.get('Kernel Name', ('?',))[0]}`\n")
    print("| metric | value |\n|---|---|")
    for k, label in KEYS:
        if k in d:
            print(f"| {label} (`{k}`) | {d[k][0]} {d[k][1]} |")
    src = page(rep, "source")
    hdr, data = src[1], src[2:]
    ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    tot, ts = 0, 0
    byop, sampop = collections.Counter(), collections.Counter()
    for r in data:
        n = int(r[ia]); tot += n; ts += int(r[isamp])
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        op = m.group(2).split(".")[0] if m else "?"
        byop[op] += n; sampop[op] += int(r[isamp])
    print(f"\nSASS: {len(data)} static instructions, {tot} executed warp instructions, {ts} stall samples.\n")
    print("| opcode | share of executed | share of samples |\n|---|---|---|")
    for op, n in byop.most_common(12):
        print(f"| {op} | {100 * n / tot:.1f} % | {100 * sampop[op] / max(ts, 1):.1f} % |")
    fp64 = sum(byop[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
    print(f"\nFP64 arithmetic share of issued instructions: {100 * fp64 / tot:.1f} %")
    tma = [r[isrc] for r in data if "UBLKCP" in r[isrc] or "SYNCS" in r[isrc]]
    print(f"TMA bulk-copy / mbarrier instructions present: {len(tma)} (UBLKCP / SYNCS)")


if __name__ == "__main__":
    main()
