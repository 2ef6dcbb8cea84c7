"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and one
denoiser evaluation's launch sequence.  Usage: python tools/launch_list_summary.py launches.csv [--seq]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    seq = []
    for row in r:
        name = row[ki].split("(")[0].replace("void ", "").replace("gc::<unnamed>::", "").replace("unnamed>::", "")
        seq.append((name, float(row[vi].replace(",", "")) / 1e3, row[gi]))
    idx = [i for i, (n, _, _) in enumerate(seq) if "dpm_update" in n]
    one = seq[idx[0] + 1: idx[1] + 1] if len(idx) >= 2 else seq
    tot = sum(v for _, v, _ in one)
    print(f"one denoiser evaluation (+ solver update): {len(one)} launches, {tot:.1f} us serialised")
    agg = collections.OrderedDict()
    for n, v, g in one:
        d = agg.setdefault(n, [0, 0.0])
        d[0] += 1
        d[1] += v
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k[:58]:58s} n={n:4d} total={t:8.1f} us avg={t / n:7.2f} us share={t / tot:.3f}")
    if "--seq" in sys.argv:
        for n, v, g in one:
            print(f"    {n[:50]:50s} {v:8.2f} us grid={g}")


if __name__ == "__main__":
    main()
