"""Per-source-line summary of an ncu capture: joins `ncu --page source --csv` (SASS-level samples / executed
instructions) with the line table of the object file (`nvdisasm -g`).
usage: python profiles/ncu_lines.py <report.ncu-rep> <object .o or .so> <kernel name substring> [top N]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) > 5]
ia, isrc, ins, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = min(int(r[ia], 16) for r in data)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
line_of = {}
for cub in os.listdir(tmp):
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    cur, inside = None, False
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and cur:
            line_of[int(m.group(1), 16)] = cur
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
tot_s = tot_i = 0
for r in data:
    key = line_of.get(int(r[ia], 16) - base, ("?", 0))
    s, e = int(r[ins]), int(r[iex])
    agg[key][0] += s
    agg[key][1] += e
    for h, i in stall_cols:
        if r[i] not in ("", "0"):
            agg[key][2][h] += int(r[i])
    tot_s += s
    tot_i += e
src_cache = {}
def src(f, l):
    if f not in src_cache:
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "eqtlbma_b200", "csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    t = src_cache[f]
    return t[l - 1].strip()[:90] if 0 < l <= len(t) else ""
print(f"samples {tot_s}  warp instructions {tot_i}")
print("%-22s %7s %7s  %-28s %s" % ("line", "samp%", "inst%", "top stalls", "source"))
for key, (s, e, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    ts = ",".join(f"{h[6:]}:{100 * v // max(s, 1)}" for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print("%-22s %6.2f%% %6.2f%%  %-28s %s" % (f"{key[0]}:{key[1]}", 100.0 * s / tot_s, 100.0 * e / tot_i, ts, src(*key)))
