"""Per-source-line executed-instruction counts for one kernel: pairs the SASS page of an ncu report (per-instruction
counts, in address order) with nvdisasm -g line markers of the cubin the report was taken from.
usage: ncu_lines.py report.ncu-rep cubin kernel_substring [launch_index]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, cubin, kname = sys.argv[1:4]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# find the function section
start = None
for i, l in enumerate(dis):
    if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"):
        start = i
        break
if start is None:
    for i, l in enumerate(dis):
        if re.match(r"\s*\.section\s+\.text\..*" + re.escape(kname), l):
            start = i
            break
assert start is not None, "kernel not found in cubin"
line_of = []  # per instruction, (file, line)
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
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kname, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hdr_i]
ie, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
insts = []
for r in rows[hdr_i + 1:]:
    if len(r) <= ie:
        continue
    if r[0].startswith("Kernel Name") or r[0] == "Address":   # next launch of the same kernel
        break
    try:
        insts.append((r[isrc].strip(), int(r[ie]), int(r[isamp])))
    except ValueError:
        pass
print(f"{len(insts)} SASS instructions in report, {len(line_of)} in cubin")
n = min(len(insts), len(line_of))
by = collections.Counter()
sm = collections.Counter()
tot = sum(c for _, c, _ in insts)
ts = sum(s for _, _, s in insts) or 1
for (src, c, s), ln in zip(insts[:n], line_of[:n]):
    by[ln] += c
    sm[ln] += s
print(f"total warp-instructions {tot}")
for ln, c in by.most_common(45):
    print(f"{ln[0]}:{ln[1]:<5d} {c / tot * 100:5.1f}% inst   {sm[ln] / ts * 100:5.1f}% samples")
