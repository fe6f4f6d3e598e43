"""Per source line (file:line, summed over all functions it was inlined into) instruction / sample shares
of one kernel: python tools/ncu_lines2.py src.csv lib.so kernel [topN]   (see ncu_lines.py)"""
import csv, re, subprocess, sys, os, tempfile, collections
src_csv, so, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
in_sec = False; line = None; fmap = {}
for ln in dis:
    if ln.startswith("//--------------------- .text."):
        in_sec = kernel in ln
        continue
    if not in_sec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        fmap[int(m.group(1), 16)] = (line, m.group(2).strip().split()[0])
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
data = rows[2:]; base = int(data[0][0], 16)
byl = collections.Counter(); sl = collections.Counter(); thr = collections.Counter(); byop = collections.Counter()
ti = ts = 0
for r in data:
    off = int(r[0], 16) - base
    ie = int(r[ci["Instructions Executed"]] or 0); ss = int(r[ci["# Samples"]] or 0); te = int(r[ci["Thread Instructions Executed"]] or 0)
    l, op = fmap.get(off, (None, "?"))
    byl[l] += ie; sl[l] += ss; thr[l] += te; ti += ie; ts += ss; byop[op.split('.')[0]] += ie
print("top lines by instructions")
for l, v in byl.most_common(top):
    print("%-26s %6.2f%% inst %6.2f%% samples  thr %.1f" % ("%s:%s" % l if l else "?", 100 * v / ti, 100 * sl[l] / max(ts, 1), thr[l] / max(v, 1)))
print("\nopcodes:", ", ".join("%s %.1f%%" % (k, 100 * v / ti) for k, v in byop.most_common(25)))
