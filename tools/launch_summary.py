"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one bench step, per kernel."""
import collections
import csv
import sys


def main(path, detail=False):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = [(r[ki], float(r[vi].replace(",", ""))) for r in rows if len(r) > vi and r[0].isdigit()]
    idx = [i for i, (k, _) in enumerate(seq) if "preprocess" in k]
    s, e = (idx[-2], idx[-1]) if len(idx) >= 2 else (0, len(seq))
    step = seq[s:e]
    if detail:
        for k, v in step:
            print(f"{v/1e3:10.1f} us  {k[:90]}")
    agg = collections.OrderedDict()
    for k, v in step:
        a = agg.setdefault(k, [0.0, 0])
        a[0] += v
        a[1] += 1
    tot = sum(v[0] for v in agg.values())
    print(f"one step: {len(step)} launches, {tot/1e6:.3f} ms (serialised, cold-cache ncu times)")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"{v[1]:4d} x {v[0]/v[1]/1e3:9.1f} us = {v[0]/1e6:8.3f} ms {v[0]/tot*100:5.1f}%  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], "--detail" in sys.argv)
