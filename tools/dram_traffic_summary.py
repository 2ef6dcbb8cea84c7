"""Per-kernel-family DRAM traffic of ONE denoiser evaluation from an ncu CSV of
`--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` over `bench.py` (launches between two
dpm_update launches = one evaluation).  Writes the JSON `bench.py` reads for `roofline.traffic`.
Usage: python tools/dram_traffic_summary.py <csv> <out.json> [workload] [members]"""
import csv
import json
import sys

FAMILIES = ["gemm_bf16_tcgen05", "ln_cond_segment_sum", "ln_cond", "edge_hidden", "khop_attention_gather", "edge_mlp_sum3",
            "dpm_update"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main():
    path, out = sys.argv[1], sys.argv[2]
    workload = sys.argv[3] if len(sys.argv) > 3 else "1deg"
    members = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    launches = {}
    for r in rows:
        d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT[r["Metric Unit"]]
    seq = [launches[i] for i in sorted(launches)]
    marks = [i for i, l in enumerate(seq) if "dpm_update" in l["name"]]
    if len(marks) < 2:
        raise SystemExit("need two dpm_update launches in the capture")
    one = seq[marks[0] + 1: marks[1] + 1]
    fam = {}
    for l in one:
        name = next((f for f in FAMILIES if f in l["name"]), None)
        if name == "edge_mlp_sum3" and ", 2>" in l["name"]:
            name = "edge_mlp_rows"          # MODE 2 of the same kernel template = gc_edge_mlp_rows
        if name is None:
            continue
        f = fam.setdefault(name, {"launches": 0, "dram_bytes_read": 0, "dram_bytes_written": 0, "us": 0.0})
        f["launches"] += 1
        f["dram_bytes_read"] += int(l["dram__bytes_read.sum"])
        f["dram_bytes_written"] += int(l["dram__bytes_write.sum"])
        f["us"] += l["gpu__time_duration.sum"]
    for f in fam.values():
        f["dram_bytes_per_launch"] = int((f["dram_bytes_read"] + f["dram_bytes_written"]) / f["launches"])
        f["us"] = round(f["us"], 1)
    doc = {"workload": workload, "members_per_gpu": members,
           "command": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                      "--cache-control none -s 1000 -c 300 --csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary",
           "csv": path,
           "note": "one denoiser evaluation of %d members (between two dpm_update launches); per kernel family: launches, "
                   "DRAM bytes read / written, serialised time" % members,
           "families": fam}
    json.dump(doc, open(out, "w"), indent=1)
    for k, v in fam.items():
        print(f"{k:24s} n={v['launches']:3d} read {v['dram_bytes_read'] / v['launches'] / 1e6:8.1f} MB  written "
              f"{v['dram_bytes_written'] / v['launches'] / 1e6:8.1f} MB per launch, {v['us'] / v['launches']:7.1f} us")


if __name__ == "__main__":
    main()
