"""Summarise `ncu -i X.ncu-rep --page source --csv` for one launch: stall samples per source line
and the hottest SASS instructions.  Usage: python tools/ncu_source_summary.py report.ncu-rep <launch-skip> [top]"""
import csv
import io
import subprocess
import sys


def main():
    rep, skip = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Line No"'))
    print("\n".join(lines[:start][:3]))
    r = csv.reader(io.StringIO("\n".join(lines[start:])))
    hdr = next(r)
    si = hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    src_rows, sass_rows = [], []
    cur = None
    for row in r:
        if len(row) < len(hdr):
            continue
        try:
            samples = int(row[si])
        except ValueError:
            continue
        if row[0] != "":
            cur = (row[0], row[1].strip()[:110])
            src_rows.append((samples, cur, row))
        else:
            sass_rows.append((samples, row[3].strip()[:90], cur, row))
    tot = sum(s for s, _, _ in src_rows) or 1
    print(f"total samples {tot}")
    print("--- hottest source lines")
    for s, (ln, txt), row in sorted(src_rows, key=lambda t: -t[0])[:top]:
        stalls = sorted(((int(row[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
        print(f"{100*s/tot:5.1f}%  L{ln:>4} {txt}   {[(n, v) for v, n in stalls if v]}")
    print("--- hottest SASS")
    for s, txt, cur, row in sorted(sass_rows, key=lambda t: -t[0])[:top]:
        stalls = sorted(((int(row[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"{100*s/tot:5.1f}%  {txt}   <- L{cur[0] if cur else '?'}  {[(n, v) for v, n in stalls if v]}")


if __name__ == "__main__":
    main()
