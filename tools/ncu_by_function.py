"""Aggregate an `ncu --page source --csv` SASS listing by device function (developer tool).
usage: python tools/ncu_by_function.py <rep.ncu-rep> <lib.so> <kernel-substring>"""
import csv, io, re, subprocess, sys, tempfile, os
rep, lib, kern = sys.argv[1:4]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if "sm_100" in f][0]
sym = subprocess.run(["readelf", "-sW", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
funcs = []
for l in sym.splitlines():
    p = l.split()
    if len(p) >= 8 and p[3] == "FUNC" and kern in p[7]:
        size = int(p[2], 0)
        name = p[7].split("$")[-1] if "$" in p[7] else "<kernel body>"
        funcs.append((int(p[1], 16), size, name))
base = 0
funcs.sort()
a0 = int(rows[0]["Address"], 16) if rows[0]["Address"].startswith("0x") else int(rows[0]["Address"])
agg = {}
tot_i = tot_s = 0
for r in rows:
    a = (int(r["Address"], 16) if r["Address"].startswith("0x") else int(r["Address"])) - a0 + base
    name = "?"
    for off, size, n in funcs:
        if n != "<kernel body>" and off <= a < off + size:
            name = n
            break
    else:
        name = "<kernel body>"
    ie = int(r["Instructions Executed"] or 0); sm = int(r["# Samples"] or 0)
    d = agg.setdefault(name, [0, 0, 0])
    d[0] += ie; d[1] += sm; d[2] += 1
    tot_i += ie; tot_s += sm
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
print("%6s %6s %7s  %s" % ("%inst", "%smpl", "n_sass", "function"))
for n, d in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6.1f %6.1f %7d  %s" % (100.0 * d[0] / tot_i, 100.0 * d[1] / max(tot_s, 1), d[2], re.sub(r"^_Z\d+", "", n)[:60]))
