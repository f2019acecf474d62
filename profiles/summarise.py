#!/usr/bin/env python
"""Summarise ncu outputs into the small text files committed under profiles/.

    python profiles/summarise.py launches gpurun_out/launches.csv            # per-kernel time shares (library kernels)
    python profiles/summarise.py full gpurun_out/prof.ncu-rep [kernel-regex] # key metrics of a --set full capture
    python profiles/summarise.py hot gpurun_out/prof.ncu-rep                 # where the instructions go (SASS page, 512-byte buckets)
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"(mlrn::)?<?unnamed>::", "", name)          # also inside template arguments (k_fit_stage<BazinPolicyT<0>, 0>)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(v for _, v in agg.values())
    lib = sum(v for k, (_, v) in agg.items() if k.startswith("k_"))
    print(f"# {path}: {sum(c for c, _ in agg.values())} launches, {tot/1e6:.3f} ms total device time "
          f"(cold-cache, serialised under ncu: compare SHARES); library kernels {lib/1e6:.3f} ms")
    print(f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share of library':>17s}")
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        if k.startswith("k_"):
            print(f"{k[:60]:60s} {c:8d} {v/1e6:10.3f} {100*v/lib:16.1f}%")
    other = [(k, c, v) for k, (c, v) in agg.items() if not k.startswith("k_")]
    print(f"{'(torch kernels: generator, copies)':60s} {sum(c for _, c, _ in other):8d} {sum(v for _, _, v in other)/1e6:10.3f}")


def full(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pattern and not re.search(pattern, r[ki]):
            continue
        print(f"## {re.sub(r'^void ', '', r[ki])[:110]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:90s} {r[i]:>18s} {units[i]}")
        print()


def hot(path, bucket=0x200, top=24):
    """Instruction and stall-sample shares per code-address bucket of the first kernel of a --set full capture
    (needs --import-source on / -lineinfo only for the source column; the SASS page is always there)."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ia, ie, isamp, ith = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
    isrc = hdr.index("Source")
    data, base = [], None
    for r in rows[2:]:
        try:
            a = int(r[ia], 16)
        except (ValueError, IndexError):
            continue
        base = a if base is None else base
        data.append((a - base, r[isrc].strip(), int(r[ie]), int(r[isamp]), float(r[ith])))
    tot, tots = sum(d[2] for d in data), sum(d[3] for d in data)
    print(f"# {rows[0][1][:100]}")
    print(f"# {len(data)} SASS instructions, {tot:.4g} warp instructions executed, {tots} stall samples")
    agg = collections.OrderedDict()
    for off, src, n, sm, th in data:
        x = agg.setdefault(off // bucket, [0, 0, 0.0, collections.Counter()])
        x[0] += n; x[1] += sm; x[2] += n * th
        x[3][src.split()[0] if src else "?"] += n
    print(f"{'code offset':>12s} {'instr %':>8s} {'samples %':>10s} {'avg threads':>12s}  dominant opcodes")
    for k, (n, sm, th, ops) in sorted(sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]):
        names = ", ".join(f"{o} {100 * c / max(n, 1):.0f}%" for o, c in ops.most_common(4))
        print(f"{hex(k * bucket):>12s} {100 * n / tot:8.1f} {100 * sm / max(tots, 1):10.1f} {th / max(n, 1):12.1f}  {names}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "hot":
        hot(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
