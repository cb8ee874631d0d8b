#!/usr/bin/env python
"""Per-source-line stall samples / executed instructions of one kernel in an .ncu-rep.
Usage: python tools/ncu_lines.py REP KERNEL_SUBSTRING [top]"""
import csv, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0      # skip this many matching blocks (one block per source file)
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
blocks = txt.split('"File Path",')
for blk in blocks[1:]:
    lines = blk.split("\n")
    fname = lines[1]
    if pat not in fname:
        continue
    if skip > 0:
        skip -= 1
        continue
    rdr = csv.reader(lines[2:])
    h = next(rdr)
    iL, iS, iW, iE = h.index("Line No"), 1, h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    rows = []

    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    for r in rdr:
        if len(r) <= iE or not r[iL]:
            continue
        st = {c: num(r[i]) for i, c in stall_cols}
        rows.append((num(r[iW]), num(r[iE]), r[iL], r[iS].strip()[:100], st))
    tot_s, tot_e = sum(r[0] for r in rows), sum(r[1] for r in rows)
    print(fname[:120]); print("total samples", tot_s, "instructions", tot_e)
    agg = collections.Counter()
    for r in rows:
        for c, v in r[4].items():
            agg[c] += v
    print("stall reasons:", ", ".join("%s %.1f%%" % (c[6:], 100. * v / max(1, tot_s)) for c, v in agg.most_common(8)))
    for w, e, ln, src, st in sorted(rows, reverse=True)[:top]:
        main = max(st.items(), key=lambda kv: kv[1])
        print("%5.1f%% samp %5.1f%% inst  L%-4s %-100s [%s]" % (100. * w / tot_s, 100. * e / tot_e, ln, src, main[0][6:]))
    break
