#!/usr/bin/env python3
"""Per-kernel summary of an .ncu-rep with source info: duration, issue utilisation, stall mix,
instruction mix and the hottest SASS lines.  Usage: python tools/ncu_mix.py file.ncu-rep [top]"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h = raw[0]
col = {k: i for i, k in enumerate(h)}
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
blocks, cur = [], None
for r in src:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r and r[0].startswith("0x"):
        cur["rows"].append(r)
# the source page repeats every kernel once per "view"; keep one block per launch
per = len(blocks) // max(1, len(raw) - 2)
blocks = blocks[::max(1, per)]
for k, row in enumerate(raw[2:]):
    g = lambda name: row[col[name]] if name in col else "?"
    print(f"=== launch {k}: {g('Kernel Name')[:60]}  grid {g('launch__grid_size')}  {g('gpu__time_duration.sum')} {raw[1][col['gpu__time_duration.sum']]}")
    print(f"    tensor pipe {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')}%  issue active {g('smsp__issue_active.avg.pct_of_peak_sustained_active')}%  "
          f"dram rd {g('dram__bytes_read.sum')} {raw[1][col['dram__bytes_read.sum']]} wr {g('dram__bytes_write.sum')} {raw[1][col['dram__bytes_write.sum']]}  "
          f"dram% {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')}  regs {g('launch__registers_per_thread')}")
    stalls = {n.replace("smsp__pcsamp_warps_issue_stalled_", ""): int(float(row[i])) for n, i in col.items()
              if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("not_issued") and row[i] not in ("", "n/a")}
    tot = sum(stalls.values()) or 1
    print("    stalls: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for n, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]))
    if k >= len(blocks):
        continue
    b = blocks[k]
    hd = b["hdr"]
    si, so, ie = hd.index("# Samples"), hd.index("Source"), hd.index("Instructions Executed")
    ti = sum(int(r[ie]) for r in b["rows"]) or 1
    ts = sum(int(r[si]) for r in b["rows"]) or 1
    mix, smp = collections.Counter(), collections.Counter()
    for r in b["rows"]:
        op = re.sub(r"^@!?U?P\d+\s+", "", r[so].strip()).split()[0].split(".")[0]
        mix[op] += int(r[ie])
        smp[op] += int(r[si])
    print(f"    warp instr {ti / 1e6:.1f}M: " + ", ".join(f"{op} {100 * c / ti:.0f}%/{100 * smp[op] / ts:.0f}%s" for op, c in mix.most_common(14)))
    for r in sorted(b["rows"], key=lambda r: -int(r[si]))[:top]:
        print(f"      {100 * int(r[si]) / ts:5.1f}%  x{int(r[ie]):>9}  {r[so].strip()[:90]}")
