#!/usr/bin/env python
"""In-order issue model of the largest basic block of each solver kernel (no GPU needed).

usage: python profiles/tools/hot_block_model.py [path/to/librsfm.so] [kernel-name-substring]

The fast interval of the solver is ONE ~1,000-instruction basic block that a warp executes once per output interval;
with one warp per sub-partition (cfg 2) its latency is the kernel time.  ptxas' schedule of that block changes with
whatever else is compiled into the kernel, so every change to the slow paths is checked here before it goes to the
GPU: the block is taken from `cuobjdump -xelf` + `nvdisasm`, and issued in order with per-register ready times
(FP64: 8 cycles result latency, 2 issue cycles per warp instruction; other ALU 5; loads 28).  Crude, but it tracked
the measurements of round 1: 1,900 vs 2,370 model cycles for two builds that ran cfg 2 in 39.2 and 48.9 ms; 2,206 vs
2,072 for the 131,072-chain kernel that went from 32.2 to 32.6 M solves/s.  Also printed: local-memory instructions
inside the block (spills in the hot path)."""
import os
import re
import subprocess
import sys
import tempfile


def blocks_of(sass_lines, kernel):
    start = [i for i, l in enumerate(sass_lines) if l.startswith(".text." + kernel + ":")][0]
    end = [i for i, l in enumerate(sass_lines) if l.startswith("//--------------------- .text.") and i > start]
    end = end[0] if end else len(sass_lines)
    blocks, cur = [], []
    for l in sass_lines[start:end]:
        if re.match(r"^\.L_x_\d+:", l):
            blocks.append(cur)
            cur = []
            continue
        m = re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s+(.*?)\s*;", l)
        if m:
            cur.append(m.group(1))
    blocks.append(cur)
    return blocks


def regs_of(tok):
    return [(m.group(1), int(m.group(2))) for m in re.finditer(r"\b(U?R)(\d+)(\.64)?\b", tok)]


def model_cycles(block, dlat=8, olat=5, ldlat=28):
    ready, t = {}, 0
    for ins in block:
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins.strip())
        parts = ins.split(None, 1)
        op, args = parts[0], (parts[1] if len(parts) > 1 else "")
        ops = [a.strip() for a in args.split(",")]
        is64 = op[0] == "D" or ".64" in op
        no_dst = op.startswith(("ST", "DSETP", "ISETP", "BRA", "BSSY", "BSYNC", "WARPSYNC", "NOP", "PLOP3", "VOTE", "BAR", "SYNCS"))
        srcs = ops if no_dst else ops[1:]
        dsts = [] if no_dst else regs_of(ops[0])
        for s in srcs:
            for b, r in regs_of(s):
                for rr in ((r, r + 1) if is64 and b == "R" else (r,)):
                    t = max(t, ready.get((b, rr), 0))
        lat = dlat if op[0] == "D" else (ldlat if op.startswith(("LDS", "LDL", "LDC", "LDG")) else olat)
        for b, r in dsts:
            for rr in ((r, r + 1) if is64 and b == "R" else (r,)):
                ready[(b, rr)] = t + lat
        t += 2 if op[0] == "D" else 1
    return t


def main():
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    lib = (sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] else
           os.path.join(root, "bayesian-markov-chain-monte-carlo_b200", "librsfm.so"))
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        sass = subprocess.run(["nvdisasm", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
    kernels = [l[len(".text."):-1] for l in sass if re.match(r"^\.text\._Z\d+rsf_(mcmc|forward|init)\w*:$", l)]
    print(f"{'kernel':72s} {'instr':>6s} {'cycles':>7s} {'LDL/STL':>8s}")
    for k in kernels:
        if want not in k:
            continue
        hot = max(blocks_of(sass, k), key=len)
        nloc = sum(1 for i in hot if re.search(r"\b(LDL|STL)", i))
        print(f"{k[:72]:72s} {len(hot):6d} {model_cycles(hot):7d} {nloc:8d}")


if __name__ == "__main__":
    main()
