#!/usr/bin/env python
"""Turn the scratch ncu output under gpurun_out/ into the small tracked summaries under profiles/.
usage: make_profiles.py <round-tag> <launches.csv> <full.ncu-rep> <workload> <variant>"""
import collections, csv, json, os, subprocess, sys

tag, launches, rep, workload, variant = sys.argv[1:6]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list -------------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(launches)) if r and r[0].isdigit()]
agg = collections.OrderedDict()
lines = []
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("unnamed>::", "").replace("hb::<", "")
    v = float(r[-1].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(r[-2], v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    lines.append("%4s  %-40s grid %-14s block %-12s %10.1f us" % (r[0], name[:40], r[8], r[7], v))
tot = sum(a[1] for a in agg.values())
with open(os.path.join(out, "%s_launches.txt" % tag), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare shares)\n")
    f.write("# command: python bench.py --steps 3 --warmup 3 --no-cpu   (device-resident steps, then the host-buffer e2e path\n")
    f.write("#          whose chunks (16 MiB; 32 MiB before r1d) are separate launches of the same kernel)\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-40s launches %3d  total %10.1f us  share %5.1f%%  avg %9.1f us\n" % (k[:40], n, t, 100 * t / tot, t / n))
    f.write("\n" + "\n".join(lines) + "\n")

# ---- full capture of the encode kernel -------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, vals = rr[0], rr[1], rr[2]
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
keep += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
d = {}
with open(os.path.join(out, "%s_encode_%s_ncu.txt" % (tag, workload)), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:encode_kernel (one launch, %s, %s)\n" % (workload, variant))
    for i, h in enumerate(hdr):
        if h in keep:
            f.write("%-95s %-18s %s\n" % (h, units[i], vals[i]))
            d[h] = (units[i], vals[i])

def to_bytes(unit, v):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]

traffic = to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
tj = os.path.join(out, "traffic.json")
t = json.load(open(tj)) if os.path.exists(tj) else {}
t[workload] = {"dram_bytes_per_launch": traffic, "kernel_variant": variant, "source": "%s_encode_%s_ncu.txt" % (tag, workload),
               "dram_read": to_bytes(*d["dram__bytes_read.sum"]), "dram_write": to_bytes(*d["dram__bytes_write.sum"])}
json.dump(t, open(tj, "w"), indent=1, sort_keys=True)
print("traffic", workload, traffic)
