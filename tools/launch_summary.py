"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel for ONE step of bench.py.

    python tools/launch_summary.py gpurun_out/launches.csv [step_index_from_end]

ncu serialises the launches and runs them cold-cache: compare SHARES of the step, not absolute times."""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    H = rows[hdr]
    ki, vi, idi = H.index("Kernel Name"), H.index("Metric Value"), H.index("ID")
    out = []
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            try:
                out.append((int(r[idi]), r[ki], float(r[vi].replace(",", ""))))
            except ValueError:
                pass
    return out


def main():
    data = load(sys.argv[1])
    back = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    names = [d[1] for d in data]
    # a step starts at the occurrence-index build (side stream) or the gather kernel, whichever comes first
    starts = [i for i, n in enumerate(names) if "emb_build_keys" in n]
    if len(starts) < back + 1:
        starts = [i for i, n in enumerate(names) if "embed_senet_fwd" in n]
    a, b = starts[-back - 1], starts[-back]
    step = data[a:b]
    tot = sum(d[2] for d in step)
    agg = collections.OrderedDict()
    for _, n, v in step:
        n = re.sub(r"^void ", "", n.split("(")[0])
        n = re.sub(r"<.*", "", n)
        agg.setdefault(n, [0.0, 0])
        agg[n][0] += v
        agg[n][1] += 1
    print(f"one step: {len(step)} launches, {tot / 1e3:.1f} us (serialised, cold cache)")
    for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{v / 1e3:9.1f} us {c:3d}x {100 * v / tot:5.1f}%  {n}")


if __name__ == "__main__":
    main()
