#!/usr/bin/env python
"""Summaries of ncu captures that were exported to CSV on the GPU box (the .ncu-rep of a --set full capture of several
kernels exceeds the size gpurun brings back):

    ncu -i prof.ncu-rep --page raw --csv > raw.csv ;  ncu -i prof.ncu-rep --page source --csv --print-source sass | gzip > source.csv.gz
    python profiles/summarise_csv.py raw gpurun_out/prof_r02a_raw.csv
    python profiles/summarise_csv.py hot gpurun_out/prof_r02a_source.csv.gz [kernel-index]
"""
import collections
import csv
import gzip
import re
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__icc_request_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def raw(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h, units = rows[hdr], rows[hdr + 1]
    ki = h.index("Kernel Name")
    for n, r in enumerate(rows[hdr + 2:]):
        if len(r) <= ki:
            continue
        print(f"## [{n}] {re.sub(r'^void ', '', r[ki])[:150]}")
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print(f"{k:90s} {r[i]:>18s} {units[i]}")
        print()


def hot(path, which=0, bucket=0x200, top=28):
    op = gzip.open if path.endswith(".gz") else open
    text = op(path, "rt", errors="replace").read()
    # the source page of several kernels is a sequence of blocks, each starting with a "Kernel Name" line
    blocks = re.split(r'(?m)^"Kernel Name",', text)
    blocks = [b for b in blocks if "Address" in b]
    rows = list(csv.reader(blocks[which].splitlines()))
    name = rows[0][0] if rows and rows[0] else "?"
    hi = next(i for i, r in enumerate(rows) if "Address" in r)
    hdr = rows[hi]
    ia, ie, isamp, ith = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
    isrc = hdr.index("Source")
    data, base = [], None
    for r in rows[hi + 1:]:
        try:
            a = int(r[ia], 16)
        except (ValueError, IndexError):
            continue
        base = a if base is None else base
        data.append((a - base, r[isrc].strip(), int(r[ie]), int(r[isamp]), float(r[ith])))
    tot, tots = sum(d[2] for d in data), sum(d[3] for d in data)
    thr = sum(d[2] * d[4] for d in data) / max(tot, 1)
    print(f"# kernel {which} of {len(blocks)}: {name[:120]}")
    print(f"# {len(data)} SASS instructions, {tot:.4g} warp instructions executed, {tots} stall samples, {thr:.1f} threads active per instruction on average")
    agg = collections.OrderedDict()
    for off, src, n, sm, th in data:
        x = agg.setdefault(off // bucket, [0, 0, 0.0, collections.Counter()])
        x[0] += n; x[1] += sm; x[2] += n * th
        x[3][src.split()[0] if src else "?"] += n
    print(f"{'code offset':>12s} {'instr %':>8s} {'samples %':>10s} {'avg threads':>12s}  dominant opcodes")
    for k, (n, sm, th, ops) in sorted(sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]):
        names = ", ".join(f"{o} {100 * c / max(n, 1):.0f}%" for o, c in ops.most_common(4))
        print(f"{hex(k * bucket):>12s} {100 * n / tot:8.1f} {100 * sm / max(tots, 1):10.1f} {th / max(n, 1):12.1f}  {names}")
    ops = collections.Counter()
    for off, src, n, sm, th in data:
        ops[src.split()[0] if src else "?"] += n
    print("# opcode mix:", ", ".join(f"{o} {100 * c / tot:.1f}%" for o, c in ops.most_common(14)))


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2])
    else:
        hot(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
