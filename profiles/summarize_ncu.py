"""Turns gpurun_out/prof_*.ncu-rep (ncu --set full) into one compact CSV + markdown table under profiles/.

    python profiles/summarize_ncu.py r1        # reads gpurun_out/, writes profiles/ncu_summary_r1.csv/.md, ncu_traffic.json
"""
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
WANT = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_read"),
        ("dram__bytes_write.sum", "dram_write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct")]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


rows_out = []
pattern = "prof_%s_*.ncu-rep" % tag if glob.glob(os.path.join(ROOT, "gpurun_out", "prof_%s_*.ncu-rep" % tag)) else "prof_*.ncu-rep"
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", pattern))):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {"capture": os.path.basename(rep).replace(".ncu-rep", "")}
        for name, short in WANT:
            if name in hdr:
                i = hdr.index(name)
                if short in ("dram_read", "dram_write"):
                    d[short] = to_bytes(r[i], units[i])
                elif short == "time_us":
                    v = float(r[i].replace(",", ""))
                    d[short] = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(units[i], 1)
                else:
                    d[short] = r[i]
        d["kernel"] = d["kernel"].replace("void ", "")[:70]
        d["dram_total"] = d.get("dram_read", 0) + d.get("dram_write", 0)
        d["dram_GBps"] = d["dram_total"] / (d["time_us"] * 1e-6) / 1e9
        rows_out.append(d)

cols = ["capture", "kernel", "time_us", "dram_read", "dram_write", "dram_total", "dram_GBps", "dram_pct_of_peak", "achieved_occupancy_pct",
        "regs", "grid", "block", "l2_hit_pct", "l1_hit_pct", "lsu_pipe_pct", "sm_pct"]
with open(os.path.join(ROOT, "profiles", "ncu_summary_%s.csv" % tag), "w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=cols, extrasaction="ignore")
    w.writeheader()
    for d in rows_out:
        w.writerow(d)
with open(os.path.join(ROOT, "profiles", "ncu_summary_%s.md" % tag), "w") as f:
    f.write("| capture | kernel | time (us) | DRAM read+write (MB) | DRAM GB/s | DRAM %% of ncu peak | occupancy %% | regs | grid | L2 hit %% |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for d in rows_out:
        f.write("| %s | `%s` | %.1f | %.1f | %.0f | %s | %s | %s | %s | %s |\n" % (
            d["capture"], d["kernel"], d["time_us"], d["dram_total"] / 1e6, d["dram_GBps"], d.get("dram_pct_of_peak", "")[:5],
            d.get("achieved_occupancy_pct", "")[:5], d.get("regs", ""), d.get("grid", ""), d.get("l2_hit_pct", "")[:5]))

# per-launch DRAM traffic of the two headline kernels, read by bench.py (roofline.traffic)
traffic = {}
for d in rows_out:
    if d["capture"].endswith("spmv") and "csr_stream_kernel<EpiAxpby" in d["kernel"]:
        traffic["csr_spmv_256"] = d["dram_total"]
    if d["capture"].endswith("spmv_split") and "csr_stream_kernel<EpiAxpby" in d["kernel"]:
        traffic["csr_spmv_256_split"] = d["dram_total"]      # the SPLIT instantiation of the row-partitioned runs, captured at world 1
    if d["capture"].endswith("spmv") and "sell_kernel<EpiAxpby" in d["kernel"]:
        traffic["sell_spmv_256"] = d["dram_total"]
# merge: a pass that captured only some kernels must not drop the other keys
tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
try:
    old = json.load(open(tp))
except Exception:
    old = {}
old.update(traffic)
json.dump(old, open(tp, "w"), indent=1)
print(open(os.path.join(ROOT, "profiles", "ncu_summary_%s.md" % tag)).read())
