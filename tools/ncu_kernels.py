#!/usr/bin/env python
"""Per-launch counters of every kernel in an .ncu-rep as JSON (read here, no GPU needed): what bench.py's
`roofline` record is computed from, next to the live CUDA-event times.
Usage: python tools/ncu_kernels.py gpurun_out/prof.ncu-rep "command that was profiled" > profiles/r02_kernels.json"""
import csv
import io
import json
import re
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "ncu_duration_ns",
    "launch__grid_size": "grid", "launch__block_size": "block", "launch__registers_per_thread": "regs",
    "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
    "smsp__inst_executed.sum": "warp_inst",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "sm__cycles_elapsed.max": "sm_cycles",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1., "us": 1e3, "ms": 1e6, "ns": 1., "s": 1e9,
        "usecond": 1e3, "msecond": 1e6, "nsecond": 1., "second": 1e9}


def short(name):
    m = re.search(r"(moments_kernel|shift_table_kernel<[^>]*>|table_walk_kernel<[^>]*>|ktable_kernel<[^>]*>|lazy_match_kernel<[^>]*>)", name)
    k = m.group(1) if m else name[:60]
    return k.replace("(int)", "").replace("(bool)", "").replace(" ", "")


def main(rep, cmd):
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    hdr, units = raw[0], raw[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = {"source": rep.split("/")[-1], "command": cmd, "kernels": {}}
    for r in raw[2:]:
        k = {}
        for key, label in WANT.items():
            if key in idx and r[idx[key]] not in ("", "n/a"):
                v = float(r[idx[key]].replace(",", ""))
                v *= UNIT.get(units[idx[key]], 1.)
                k[label] = v
        name = short(r[idx["Kernel Name"]])
        n = name
        c = 2
        while n in out["kernels"]:
            n = "%s#%d" % (name, c)
            c += 1
        out["kernels"][n] = k
    json.dump(out, sys.stdout, indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
