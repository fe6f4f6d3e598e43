"""Join an ncu SASS source page (CSV) with nvdisasm line info to get per-function / per-line
instruction counts and stall samples for one kernel.  Usage:
  ncu -i rep.ncu-rep --page source --csv > src.csv
  python tools/ncu_lines.py src.csv path/to/lib.so k_shipice_step [topN]"""
import csv, re, subprocess, sys, os, tempfile, collections
src_csv, so, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
in_sec = False; func = None; line = None; fmap = {}
for ln in dis:
    if ln.startswith("//--------------------- .text."):
        in_sec = kernel in ln
        continue
    if not in_sec: continue
    m = re.match(r"^(\$?[_A-Za-z0-9$]+):\s*$", ln)
    if m and not ln.startswith(".L_x"):
        name = m.group(1)
        func = name[:40]
        name = name.split("$")[-1] if "$" in name else name
        if name.startswith("_ZN"):
            # nested names: <len><id><len><id>...; keep the last component (the function)
            i = 3; parts = []
            while i < len(name) and name[i].isdigit():
                j = i
                while name[j].isdigit(): j += 1
                n = int(name[i:j]); parts.append(name[j:j + n]); i = j + n
            if parts: func = parts[-1]
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        fmap[int(m.group(1), 16)] = (func, line, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][0], 16)
byf = collections.Counter(); byl = collections.Counter(); sf = collections.Counter(); sl = collections.Counter(); thr = collections.Counter()
tot_i = tot_s = 0
for r in data:
    off = int(r[0], 16) - base
    ie = int(r[ci["Instructions Executed"]] or 0); ss = int(r[ci["# Samples"]] or 0); te = int(r[ci["Thread Instructions Executed"]] or 0)
    f, l, txt = fmap.get(off, ("?", None, ""))
    byf[f] += ie; sf[f] += ss; thr[f] += te; byl[(f, l)] += ie; sl[(f, l)] += ss
    tot_i += ie; tot_s += ss
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
print("%-22s %8s %8s %8s" % ("function", "inst%", "samples%", "avg thr"))
for f, v in byf.most_common(20):
    print("%-22s %7.2f%% %7.2f%% %8.1f" % (f, 100 * v / tot_i, 100 * sf[f] / max(tot_s, 1), thr[f] / max(v, 1)))
print("\ntop lines by samples")
for (f, l), v in sl.most_common(top):
    print("%-18s %-22s %7.2f%% samples %7.2f%% inst" % (f, "%s:%s" % l if l else "?", 100 * v / max(tot_s, 1), 100 * byl[(f, l)] / tot_i))
# optional: absolute per-line counts for one file range:  ... topN file.cuh lo hi
if len(sys.argv) > 7:
    fn, lo, hi = sys.argv[5], int(sys.argv[6]), int(sys.argv[7])
    agg = collections.Counter(); th = collections.Counter(); ns = collections.Counter()
    for r in data:
        off = int(r[0], 16) - base
        f, l, txt = fmap.get(off, ("?", None, ""))
        if l and l[0] == fn and lo <= l[1] <= hi:
            agg[l[1]] += int(r[ci["Instructions Executed"]] or 0); th[l[1]] += int(r[ci["Thread Instructions Executed"]] or 0); ns[l[1]] += 1
    for l in sorted(agg):
        print("%s:%d  warp-inst %d  (%d SASS)  per-SASS %.0f  avg thr %.1f" % (fn, l, agg[l], ns[l], agg[l] / ns[l], th[l] / max(agg[l], 1)))
# optional: top lines by instruction count inside one function:  ... topN FUNC
if len(sys.argv) == 6:
    fn = sys.argv[5]
    items = [(v, l) for (f, l), v in byl.items() if f == fn]
    items.sort(reverse=True)
    print("\ntop lines by inst in", fn)
    for v, l in items[:top]:
        print("%-24s %6.2f%% inst %6.2f%% samples" % ("%s:%s" % l if l else "?", 100 * v / tot_i, 100 * sl[(fn, l)] / max(tot_s, 1)))
