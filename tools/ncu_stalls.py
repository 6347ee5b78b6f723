"""Per-source-line stall-sample breakdown of one kernel: pairs the SASS page of an ncu report (--set full --import-source on)
with the nvdisasm -g line markers of the cubin the report was taken from.
usage: ncu_stalls.py report.ncu-rep libdynrf_b200.so kernel_substring [top_n]"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep, so, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
line_of = None
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    start = None
    for i, l in enumerate(dis):
        if re.match(r"\s*\.section\s+\.text\..*" + re.escape(kname), l) or (l.startswith(".text.") and kname in l and l.rstrip().endswith(":")):
            start = i
            break
    if start is None:
        continue
    line_of = []
    cur = ("?", 0)
    for l in dis[start + 1:]:
        if re.match(r"\s*\.section\s", l) and line_of:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            line_of.append(cur)
    break
assert line_of, "kernel not found in any cubin of " + so
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hdr_i]
ie, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
insts = []
for r in rows[hdr_i + 1:]:
    if len(r) <= ie or r[0].startswith("Kernel Name") or r[0] == "Address":
        if insts:
            break
        continue
    try:
        insts.append((r[isrc].strip(), int(r[ie]), int(r[isamp]), {n: int(r[i]) for i, n in stall_cols if r[i] not in ("", "0")}))
    except ValueError:
        pass
print(f"{len(insts)} SASS instructions in report, {len(line_of)} in cubin")
n = min(len(insts), len(line_of))
by, sm = collections.Counter(), collections.Counter()
st = collections.defaultdict(collections.Counter)
hot = collections.defaultdict(collections.Counter)
tot = sum(c for _, c, _, _ in insts) or 1
ts = sum(s for _, _, s, _ in insts) or 1
allst = collections.Counter()
for (src, c, s, d), ln in zip(insts[:n], line_of[:n]):
    by[ln] += c
    sm[ln] += s
    for k, v in d.items():
        st[ln][k] += v
        allst[k] += v
    hot[ln][src.split()[0] if src else "?"] += s
print("total warp-instructions", tot, "samples", ts)
print("stall mix:", {k: round(100 * v / ts, 1) for k, v in allst.most_common(8)})
for ln, s in sm.most_common(topn):
    top = ", ".join(f"{k} {100 * v / ts:.1f}" for k, v in st[ln].most_common(3))
    ops = ",".join(k for k, _ in hot[ln].most_common(2))
    print(f"{ln[0]}:{ln[1]:<5d} {s / ts * 100:5.1f}% samples  {by[ln] / tot * 100:5.1f}% inst   [{top}]  {ops}")
