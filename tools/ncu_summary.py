#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/ (tracked).

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python tools/ncu_summary.py raw gpurun_out/prof.ncu-rep        > profiles/rNN_prof.txt   (needs `ncu` on PATH)
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg"]
STALL = "smsp__average_warps_issue_stalled_"


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"][:110]
        ns = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms total (ncu-serialised, cold-cache: compare shares)")
    print(f"{'n':>5} {'total ms':>11} {'avg us':>11} {'share':>8}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[0]:5d} {v[1] / 1e6:11.3f} {v[1] / v[0] / 1e3:11.1f} {100 * v[1] / tot:7.2f}%  {k}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print(f"## {d.get('Kernel Name', ('', '?'))[1]}  grid {d.get('Grid Size', ('', '?'))[1]} block {d.get('Block Size', ('', '?'))[1]}")
        for k in KEYS:
            if k in d:
                print(f"{k:85s} {d[k][1]:>18s} {d[k][0]}")
        print("# warp stall reasons (warps stalled per issue-active cycle)")
        st = [(k[len(STALL):-len('_per_issue_active.ratio')], float(v[1])) for k, v in d.items()
              if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v[1]]
        for name, val in sorted(st, key=lambda kv: -kv[1]):
            print(f"    {name:28s} {val:8.3f}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] in ("launches", "raw"):
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])


def hot(path, top=40):
    """Top stall-sample SASS instructions (ncu --page source) with their dominant stall reason."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    ci = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for idx, r in enumerate(rows[start + 1:]):
        if len(r) < len(hdr) or r[0] == "Address" or r[0] == "Kernel Name":
            continue
        try:
            s = int(r[ci["# Samples"]])
        except ValueError:
            continue
        data.append((idx, s, r))
    tot = sum(s for _, s, _ in data)
    print(f"# {path}: {tot} stall samples over {len(data)} SASS instructions")
    agg = collections.Counter()
    for _, s, r in data:
        for h in stall_cols:
            agg[h] += int(r[ci[h]] or 0)
    print("# totals by reason: " + ", ".join(f"{k[6:]} {v}" for k, v in agg.most_common(8)))
    for idx, s, r in sorted(data, key=lambda x: -x[1])[:top]:
        reasons = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
        print(f"{s:7d} {100 * s / tot:5.1f}%  #{idx:5d}  {r[ci['Source']].strip()[:90]:90s}  {reasons[0][1]}:{reasons[0][0]} {reasons[1][1]}:{reasons[1][0]}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "hot":
    hot(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
