"""Per-kernel table of one train step from an `ncu --set full` capture of `bench.py --steps 1 --warmup 3 --no-graph ...`:
DRAM / L2 / L1 bytes per launch, time, utilisation of each unit by ncu's own peaks.

    python tools/kernel_table.py gpurun_out/r02_step_full.ncu-rep > profiles/r02_kernel_table.json

bench.py reads the result (roofline.traffic of the kernels it reports)."""
import collections
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from k1_table import ncu_rows  # noqa: E402


def norm(name):
    name = re.sub(r"\((bool|int)\)", "", name)
    name = re.sub(r"^(void\s+)?(dynrf::)?", "", name)
    name = re.sub(r"\(.*$", "", name)          # drop the argument list
    return name.replace(" ", "")


def main():
    rep = sys.argv[1]
    rows = ncu_rows(rep)
    agg = collections.OrderedDict()
    for r in rows:
        k = norm(r["kernel"])
        a = agg.setdefault(k, dict(launches=0, ms=0.0, dram_bytes=0.0, lts_bytes=0.0, l1tex_bytes=0.0, pct=collections.defaultdict(float)))
        a["launches"] += 1
        a["ms"] += (r["ns"] or 0) / 1e6
        a["dram_bytes"] += (r["dram_read"] or 0) + (r["dram_write"] or 0)
        a["lts_bytes"] += r["lts_bytes"] or 0
        a["l1tex_bytes"] += r["l1tex_bytes"] or 0
        for p in ("dram_pct", "lts_pct", "l1tex_pct", "sm_pct", "issue_pct", "warps_active_pct"):
            a["pct"][p] += (r[p] or 0) * ((r["ns"] or 0) / 1e6)
    out = {}
    total = sum(a["ms"] for a in agg.values())
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        n = a["launches"]
        out[k] = dict(launches=n, ms=a["ms"] / n, ms_total=a["ms"], share=a["ms"] / total if total else None,
                      dram_bytes=a["dram_bytes"] / n, lts_bytes=a["lts_bytes"] / n, l1tex_bytes=a["l1tex_bytes"] / n,
                      pct_of_peak={p[:-4]: (v / a["ms"] if a["ms"] else None) for p, v in a["pct"].items()})
    print(json.dumps({"source": f"ncu --set full --clock-control none ({os.path.basename(rep)}); per-launch means; times are under ncu "
                                "(serialised, cold cache) -- shares, not absolutes", "total_ms_under_ncu": total, "kernels": out}, indent=1))


if __name__ == "__main__":
    main()
