"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch
list: the last complete bench step (preprocess launch to preprocess launch), per kernel."""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    d = collections.OrderedDict()
    for r in rows:
        if len(r) > vi and r[0].isdigit():
            d.setdefault(int(r[idi]), {"k": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))
    return list(d.values())


def main(path, detail=False):
    seq = load(path)
    idx = [i for i, x in enumerate(seq) if "preprocess" in x["k"]]
    step = seq[idx[-2]:idx[-1]] if len(idx) >= 2 else seq
    T = "gpu__time_duration.sum"
    if detail:
        for x in step:
            print(f"{x[T]/1e3:10.1f} us  rd {x.get('dram__bytes_read.sum', 0)/1e6:8.1f} MB  wr {x.get('dram__bytes_write.sum', 0)/1e6:8.1f} MB  {x['k'][:80]}")
    agg = collections.OrderedDict()
    for x in step:
        a = agg.setdefault(x["k"], [0.0, 0, 0.0, 0.0])
        a[0] += x[T]
        a[1] += 1
        a[2] += x.get("dram__bytes_read.sum", 0.0)
        a[3] += x.get("dram__bytes_write.sum", 0.0)
    tot = sum(v[0] for v in agg.values())
    print(f"one step: {len(step)} launches, {tot/1e6:.3f} ms (serialised, cold-cache ncu times)")
    print("   n x   avg us  =  total ms  share   DRAM rd/wr MB per launch   kernel")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"{v[1]:4d} x {v[0]/v[1]/1e3:8.1f} = {v[0]/1e6:8.3f} ms {v[0]/tot*100:5.1f}%  {v[2]/v[1]/1e6:8.1f} /{v[3]/v[1]/1e6:8.1f}  {k[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], "--detail" in sys.argv)
