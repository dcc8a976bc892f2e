"""Join an ncu SASS listing (--page source --csv) with nvdisasm -g line info: instructions and stall samples per source line.
usage: python tools/ncu_by_line.py <rep.ncu-rep> <lib.so> <kernel-substring> [top N]   (developer tool)"""
import csv, io, os, re, subprocess, sys, tempfile
rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern.split("ILi")[0]], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
a0 = int(rows[0]["Address"], 16)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if "sm_100" in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
insec = False
cur = ("?", 0)
off2line = {}
for l in dis.splitlines():
    if l.startswith("//-----"):
        insec = (".text." in l and kern in l)
        continue
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
agg = {}
ti = ts = 0
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
for r in rows:
    if not re.match(r"^(0x)?[0-9a-f]+$", r["Address"] or ""): break       # the next kernel's table
    off = int(r["Address"], 16) - a0
    key = off2line.get(off, ("?", 0))
    ie = int(r["Instructions Executed"] or 0); sm = int(r["# Samples"] or 0)
    d = agg.setdefault(key, [0, 0, {}])
    d[0] += ie; d[1] += sm
    for c in stall_cols:
        v = int(r[c] or 0)
        if v: d[2][c] = d[2].get(c, 0) + v
    ti += ie; ts += sm
src = {}
def text(f, n):
    if f not in src:
        for root, _, files in os.walk(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))):
            if f in files:
                src[f] = open(os.path.join(root, f)).read().splitlines(); break
        else:
            src[f] = []
    return src[f][n - 1].strip()[:90] if 0 < n <= len(src[f]) else ""
print("total warp instructions %d, stall samples %d" % (ti, ts))
for (f, n), d in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = sorted(d[2].items(), key=lambda kv: -kv[1])[:2]
    print("%5.1f %5.1f  %s:%d  %-22s %s" % (100.0 * d[0] / ti, 100.0 * d[1] / max(ts, 1), f, n,
          ",".join("%s=%d%%" % (k[6:], 100 * v / max(d[1], 1)) for k, v in st), text(f, n)))
