"""Map ncu SASS-level stall samples to source lines via nvdisasm line info.
usage: ncu_lines.py report.ncu-rep kernel_regex mangled_substr [top [skip]]   (cubin extracted from libteb_b200.so)"""
import csv, io, subprocess, sys, collections, re, os, glob, tempfile
rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
skip = sys.argv[5] if len(sys.argv) > 5 else "0"   # matching launches to skip inside the report
by_inst = len(sys.argv) > 6 and sys.argv[6] == "inst"   # rank by executed warp instructions instead of stall samples
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "teb_local_planner_b200", "libteb_b200.so")], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
# offset -> (file,line) for the wanted function
off2line = {}
infunc = False; cur = None
for ln in dis.splitlines():
    if ln.startswith("//--------------------- .text."):
        infunc = mangled in ln
        cur = None
        continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
    if m and cur:
        off2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre, "-s", skip, "-c", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; base = None
agg = collections.Counter(); inst = collections.Counter()
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    try:
        addr = int(d["Address"], 16); samp = int(d["# Samples"]); ie = int(d["Instructions Executed"])
    except Exception:
        continue
    if base is None: base = addr
    key = off2line.get(addr - base, ("?", 0))
    agg[key] += samp; inst[key] += ie
tot = sum(agg.values()) or 1
print("total samples", tot, "total warp-instr", sum(inst.values()))
srcs = {}
order = (inst if by_inst else agg).most_common(top)
for (f, l), _ in order:
    v = agg[(f, l)]
    if f not in srcs:
        for cand in glob.glob(os.path.join(root, "**", f), recursive=True):
            srcs[f] = open(cand).read().splitlines(); break
        else:
            srcs[f] = []
    text = srcs[f][l - 1].strip()[:95] if 0 < l <= len(srcs[f]) else ""
    print(f"{100*v/tot:5.1f}%  inst {inst[(f,l)]:>9}  {f}:{l}  {text}")
