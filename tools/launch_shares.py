"""ncu launch list (--metrics gpu__time_duration.sum --csv) of a bench command -> per-kernel shares of one step.

    python tools/launch_shares.py gpurun_out/r02_launches_train.csv STEPS > profiles/r02_launch_shares_train.json

STEPS = how many steps the captured window spans (launches of the once-per-step kernel k_composite are counted to check it).
Times under ncu are serialised and cold-cache: shares, not absolutes."""
import collections
import csv
import json
import re
import sys


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path, errors="replace")))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Value" in r)
    h = rows[hdr_i]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
    seq = []
    for r in rows[hdr_i + 1:]:
        if len(r) <= mv:
            continue
        try:
            t = float(r[mv].replace(",", "")) * scale.get(r[mu], 1e-3)
        except ValueError:
            continue
        name = re.sub(r"\((bool|int)\)", "", r[kn])
        name = re.sub(r"^(void )?(dynrf::)?", "", name)
        name = re.sub(r"\(.*$", "", name).strip()
        if name.startswith("at::") or "at::native" in name:
            name = "ATen: " + re.sub(r"<.*", "", name.replace("at::native::", "").replace("at::", ""))[:60]
        seq.append((name, t))
    # whole steps only: a step ends with the optimiser kernel (k_adam_flat; round 1: the last multi_tensor_apply of torch's Adam)
    ends = [i for i, (n, _) in enumerate(seq) if n == "k_adam_flat"]
    if len(ends) >= 2:
        seq = seq[ends[0] + 1: ends[-1] + 1]
        steps = len(ends) - 1
    else:
        steps = int(sys.argv[2]) if len(sys.argv) > 2 else max(1, sum(1 for n, _ in seq if n == "k_composite"))
    cnt, tim = collections.Counter(), collections.Counter()
    for name, t in seq:
        cnt[name] += 1
        tim[name] += t
    total = sum(tim.values())
    out = [{"kernel": k, "launches_per_step": cnt[k] / steps, "us_per_step": v / steps, "share": v / total}
           for k, v in sorted(tim.items(), key=lambda kv: -kv[1])]
    aten = sum(o["us_per_step"] for o in out if o["kernel"].startswith("ATen"))
    aten_n = sum(o["launches_per_step"] for o in out if o["kernel"].startswith("ATen"))
    print(json.dumps({"source": f"ncu --metrics gpu__time_duration.sum --clock-control none ({path}); serialised, cold-cache: shares, not absolutes",
                      "steps_in_capture": steps, "us_per_step_total": total / steps, "aten_us_per_step": aten, "aten_launches_per_step": aten_n,
                      "own_launches_per_step": sum(o["launches_per_step"] for o in out) - aten_n, "kernels": out}, indent=1))


if __name__ == "__main__":
    main()
