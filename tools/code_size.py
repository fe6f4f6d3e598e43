"""Static SASS size per device function of one kernel in a .so (development aid).
usage: code_size.py lib.so kernel_name"""
import re, subprocess, os, sys, tempfile, collections
so, kernel = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
in_sec = False; func = None; byfunc = collections.Counter()
for ln in dis:
    if ln.startswith("//--------------------- .text."):
        in_sec = kernel in ln; continue
    if not in_sec: continue
    m = re.match(r"^(\$?[_A-Za-z0-9$]+):\s*$", ln)
    if m and not ln.startswith(".L_x"):
        name = m.group(1); mm = re.search(r"_ZN2bp(\d+)([a-z_0-9]+)", name)
        func = mm.group(2)[:int(mm.group(1))] if mm else name[:40]; continue
    if re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln): byfunc[func] += 1
tot = sum(byfunc.values())
print("total %d instr = %.1f KB" % (tot, tot * 16 / 1024))
for f, v in byfunc.most_common(40): print("%-34s %6d instr %6.1f KB" % (f, v, v * 16 / 1024))
