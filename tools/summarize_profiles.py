"""Summarise ncu artefacts from gpurun_out/ into profiles/ (tracked). usage: summarize_profiles.py TAG launches.csv full.ncu-rep[,second.ncu-rep,...] [B]"""
import csv, io, json, os, subprocess, sys, collections
tag, launches_csv, rep = sys.argv[1], sys.argv[2], sys.argv[3]
B = sys.argv[4] if len(sys.argv) > 4 else "?"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles")
os.makedirs(out_dir, exist_ok=True)
lines = []
# ---- launch list: time share per kernel (cold-cache, serialised: compare SHARES)
rows = [r for r in csv.reader(l for l in open(launches_csv) if l.startswith('"')) if len(r) > 5]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    name = r[ki].split("(")[0].replace("tebgpu::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
unit = rows[1][hdr.index("Metric Unit")]
lines.append(f"## launch list ({os.path.basename(launches_csv)}; ncu --metrics gpu__time_duration.sum --clock-control none; B={B})\n")
lines.append("| kernel | launches | total | share |\n|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| {k} | {c} | {t:.1f} {unit} | {100*t/tot:.1f} % |")
with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
    f.write("".join(l for l in open(launches_csv) if l.startswith('"')))
# ---- full capture: key metrics per kernel
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__thread_inst_executed_per_inst_executed.ratio"]
traffic = {}
lines.append(f"\n## full captures ({', '.join(os.path.basename(x) for x in rep.split(','))}; ncu --set full --clock-control none --import-source on)\n")
all_rows = []
for one in rep.split(","):
    raw = subprocess.run(["ncu", "-i", one, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    for r in rr[2:]:
        all_rows.append((rr[0], rr[1], r))
for h, units, r in all_rows:
    rr = [h, units]
    name = r[h.index("Kernel Name")].split("(")[0].replace("tebgpu::", "")
    lines.append(f"### {name}\n")
    vals = {}
    for w in want:
        if w in h:
            vals[w] = r[h.index(w)]
            lines.append(f"* `{w}` = {r[h.index(w)]} {rr[1][h.index(w)]}")
    st = []
    for i, hh in enumerate(h):
        if "pcsamp_warps_issue_stalled" in hh and not hh.endswith("not_issued"):
            try:
                st.append((float(r[i]), hh.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    st.sort(reverse=True)
    tots = sum(v for v, _ in st) or 1
    lines.append("* top stall reasons (pc samples): " + ", ".join(f"{n} {100*v/tots:.0f}%" for v, n in st[:5]))
    def tobytes(x, u):
        x = float(x.replace(",", ""))
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    try:
        rd = tobytes(vals["dram__bytes_read.sum"], rr[1][h.index("dram__bytes_read.sum")])
        wr = tobytes(vals["dram__bytes_write.sum"], rr[1][h.index("dram__bytes_write.sum")])
        traffic.setdefault(name, int(rd + wr))
        lines.append(f"* DRAM traffic per launch = {(rd+wr)/1e6:.1f} MB")
    except Exception:
        pass
    lines.append("")
with open(os.path.join(out_dir, f"{tag}_kernels.md"), "w") as f:
    f.write(f"# ncu summary {tag}\n\n" + "\n".join(lines) + "\n")
print("\n".join(lines))
print(json.dumps(traffic))
