#!/usr/bin/env python
"""Where the samples of an ncu capture fall, by region of the SASS listing.

usage: python profiles/tools/sass_regions.py X_source_sass.csv[.gz] [min_share_percent]

Reads `ncu -i X.ncu-rep --page source --csv --print-source sass`.  Consecutive instructions are merged into a region
while their executed count stays within 2 % (one basic block, or blocks always executed together); for every region:
first address offset, instructions, executions per instruction, share of all warp-stall samples, issue slots
(instructions executed) share, samples per execution (~ cycles the region costs a warp per pass, in sampling
periods), and the leading stall reasons."""
import csv, gzip, io, sys, collections

def main():
    f = sys.argv[1]
    min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    op = gzip.open if f.endswith(".gz") else open
    rows = list(csv.reader(io.TextIOWrapper(op(f, "rb"), errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Address" in r[:1]][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = int(data[0][ia], 16)
    tot_s = sum(int(r[isamp]) for r in data)
    tot_i = sum(int(r[iex]) for r in data)
    regions, cur = [], None
    for r in data:
        ex, s = int(r[iex]), int(r[isamp])
        if cur is None or not (0.98 * cur["ex"] <= ex <= 1.02 * cur["ex"]) or ex == 0 != (cur["ex"] == 0):
            cur = {"off": int(r[ia], 16) - base, "n": 0, "ex": max(ex, 1e-9), "s": 0, "i": 0, "st": collections.Counter(), "ops": collections.Counter()}
            regions.append(cur)
        cur["n"] += 1; cur["s"] += s; cur["i"] += ex
        cur["ops"][r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]] += 1
        for i, h in stall_cols:
            cur["st"][h] += int(r[i] or 0)
    print(f"total samples {tot_s}, warp instructions executed {tot_i}, SASS instructions {len(data)}")
    print(f"{'offset':>8s} {'instr':>6s} {'exec/instr':>12s} {'samples%':>9s} {'issue%':>7s} {'samp/exec':>10s}  leading stalls / ops")
    for g in regions:
        share = 100.0 * g["s"] / tot_s
        if share < min_share:
            continue
        st = ", ".join(f"{k[6:]} {100.0 * v / max(1, sum(g['st'].values())):.0f}%" for k, v in g["st"].most_common(3))
        ops = " ".join(f"{k}x{v}" for k, v in g["ops"].most_common(4))
        print(f"{g['off']:8x} {g['n']:6d} {g['ex']:12.0f} {share:9.2f} {100.0 * g['i'] / tot_i:7.2f} {g['s'] / g['ex']:10.5f}  {st} | {ops}")

if __name__ == "__main__":
    main()
