#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per kernel launch the headline metrics and the
executed-instruction mix.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/x.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__inst_executed_pipe_fma.sum", "  of which FMA pipe"),
    ("sm__inst_executed_pipe_fp64.sum", "  of which FP64 pipe"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "  bank-conflict wavefronts"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep):
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# %s" % rep)
    for r in raw[2:]:
        print("\n## %s" % r[idx["Kernel Name"]][:110])
        for key, label in WANT:
            if key in idx:
                print("  %-30s %s %s" % (label, r[idx[key]], units[idx[key]]))
    src = ncu(["-i", rep, "--page", "source", "--csv"])
    seen = set()
    for sec in src.split('"Kernel Name",')[1:]:
        lines = sec.split("\n")
        name = lines[0].strip().strip('",')[:110]
        if name in seen:
            continue
        seen.add(name)
        rdr = csv.reader(lines[1:])
        h = next(rdr)
        iS, iE, iW = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
        ops, samp, tot, rows = collections.Counter(), collections.Counter(), 0, []
        for r in rdr:
            if len(r) <= iE:
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
            if not m:
                continue
            op = m.group(2).split(".")[0]
            n = int(r[iE] or 0)
            ops[op] += n
            tot += n
            samp[op] += int(r[iW] or 0)
            rows.append((int(r[iW] or 0), n, r[iS].strip()[:80]))
        print("\n## instruction mix: %s\n  total warp instructions %d" % (name, tot))
        for op, n in ops.most_common(12):
            print("   %-10s %12d %5.1f%%  stall samples %d" % (op, n, 100. * n / max(tot, 1), samp[op]))
        print("  most-sampled instructions:")
        for s_, n, txt in sorted(rows, reverse=True)[:8]:
            print("   %7d samples %10d exec  %s" % (s_, n, txt))


if __name__ == "__main__":
    main(sys.argv[1])
