#!/usr/bin/env python3
"""Turn gpurun_out/launches_<tag>.csv (+ prof_*_<tag>.ncu-rep) into small tracked summaries under profiles/."""
import collections
import csv
import os
import re
import subprocess
import sys

tag = sys.argv[1]
out_dir = "profiles"
os.makedirs(out_dir, exist_ok=True)
lines = [l for l in open(f"gpurun_out/launches_{tag}.csv") if l.startswith('"')]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
    k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
with open(f"{out_dir}/launches_{tag}.txt", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {len(lines) - 1} launches of one sampler step "
            f"(python bench.py --steps 1 --warmup 3 --no-cpu-baseline); cold-cache serialised times: compare shares\n")
    f.write(f"total {tot:.1f} us\n")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:50s} n={n:4d} {v:10.1f} us {100 * v / tot:5.1f}%\n")
print(open(f"{out_dir}/launches_{tag}.txt").read())
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "launch__shared_mem_per_block_dynamic", "Block Size"]
for rep in sorted(p for p in os.listdir("gpurun_out") if p.endswith(f"_{tag}.ncu-rep")):
    raw = subprocess.run(["ncu", "-i", f"gpurun_out/{rep}", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if not rows:
        continue
    idx = [i for i, h in enumerate(rows[0]) if h in WANT]
    name = rep.replace(".ncu-rep", "")
    with open(f"{out_dir}/{name}.csv", "w") as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i] for i in idx])
    print(open(f"{out_dir}/{name}.csv").read())

# DRAM traffic of every gemm_kernel launch of one decoder forward (ncu_capture.sh: gemm_traffic_<tag>.csv)
tcsv = f"gpurun_out/gemm_traffic_{tag}.csv"
if os.path.exists(tcsv):
    import json
    rows = list(csv.DictReader(l for l in open(tcsv) if l.startswith('"')))
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r["ID"], {})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(unit, 1)
        d[r["Metric Name"]] = v
    rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in per.values())
    wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in per.values())
    ms = sum(d.get("gpu__time_duration.sum", 0.0) for d in per.values())
    out = {"gemm_kernel_dram_bytes_per_launch": (rd + wr) / max(1, len(per)), "launches": len(per),
           "dram_read_bytes_total": rd, "dram_write_bytes_total": wr, "serialized_time_ms_total": ms,
           "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm_kernel: every gemm_kernel "
                     f"launch of one decoder forward (N=128, T=1219), tools/ncu_capture.sh {tag}"}
    json.dump(out, open(f"{out_dir}/traffic_r2.json", "w"), indent=1)
    print(json.dumps(out, indent=1))
