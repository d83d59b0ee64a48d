#!/usr/bin/env python
"""Compact text summary of an `ncu --set full` capture exported with `ncu -i x.ncu-rep --page raw --csv` (and, if given,
the SASS-level `--page source --csv` export): the numbers DESIGN.md and bench.py quote, in one screen.

  python tools/ncu_summary.py gpurun_out/r2d/prof_walk.raw.csv [gpurun_out/r2d/prof_walk.source.csv] > profiles/...txt"""
import collections
import csv
import re
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"), ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__waves_per_multiprocessor", "waves/SM"), ("launch__occupancy_limit_registers", "occupancy limit: registers (blocks)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit: shared memory (blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active, % of peak"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy, %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / scheduler"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe ALU, % of peak"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe FMA, % of peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "pipe FP64, % of peak"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe XU, % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe LSU, % of peak"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "pipe uniform, % of peak"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate, %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate, %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(vals, units)))
    print("kernel:", d.get("Kernel Name", ("?",))[0])
    for k, label in KEYS:
        if k in d and d[k][0] != "":
            print("  %-44s %s %s" % (label, d[k][0], d[k][1]))
    print("  stall reasons (warps stalled per issued instruction):")
    st = []
    for h, v in zip(hdr, vals):
        m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio", h)
        if m and "not_issued" not in h and v:
            st.append((float(v), m.group(1)))
    for v, name in sorted(st, reverse=True)[:9]:
        print("      %-22s %.3f" % (name, v))
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        hdr, data = rows[1], rows[2:]
        isrc, iex, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        tot = sum(int(r[iex]) for r in data) or 1
        tst = sum(int(r[ist]) for r in data) or 1
        ops, stl = collections.Counter(), collections.Counter()
        for r in data:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
            full = m.group(2) if m else "?"
            op = full.split(".")[0]
            key = ".".join(full.split(".")[:2]) if op in ("VIADDMNMX", "VIMNMX3", "VIADD", "IMAD", "LOP3", "DFMA", "DMUL", "DADD") else op
            ops[key] += int(r[iex])
            stl[key] += int(r[ist])
        print("  executed instruction mix (SASS, %d static instructions):" % len(data))
        for k, v in ops.most_common(16):
            print("      %-18s %5.2f %% of executed, %5.2f %% of stall samples" % (k, 100.0 * v / tot, 100.0 * stl[k] / tst))


if __name__ == "__main__":
    main()
