"""Key metrics of every launch in an .ncu-rep (ncu --set full): python tools/ncu_key_metrics.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__registers_per_thread"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("---", r[hdr.index("ID")], name.split("(")[0][-60:])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:70s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main()
