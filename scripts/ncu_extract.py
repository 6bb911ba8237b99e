"""Turn the .ncu-rep files that scripts/profile_ncu.sh leaves in gpurun_out/ into the small CSV / JSON extracts that are
committed under profiles/ (run HERE, `ncu -i` needs no GPU).

    python scripts/ncu_extract.py r01final
"""
import csv
import gzip
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    return names, units, rows[hdr + 2:]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    go = os.path.join(ROOT, "gpurun_out")
    summary = {}
    for f in sorted(os.listdir(go)):
        if not (f.startswith(f"prof_{tag}_") and f.endswith(".ncu-rep")):
            continue
        names, units, launches = raw_page(os.path.join(go, f))
        kcol = names.index("Kernel Name")
        for li, row in enumerate(launches):
            label = f"{f[len('prof_' + tag) + 1:-8]}_{li}"
            dst = os.path.join(ROOT, "profiles", f"{tag}_ncu_{label}.csv")
            with open(dst, "w", newline="") as fh:
                w = csv.writer(fh)
                w.writerow(["metric", "unit", "value"])
                w.writerow(["kernel", "", row[kcol]])
                for k in KEEP:
                    if k in names:
                        w.writerow([k, units[names.index(k)], row[names.index(k)]])
            d = {k: row[names.index(k)] for k in KEEP if k in names}
            d["kernel"] = row[kcol]
            summary[label] = d
            print(label, row[kcol][:60], d.get("gpu__time_duration.sum"), d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"))
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.json"), "w") as fh:
        json.dump(summary, fh, indent=1)
    lst = os.path.join(go, f"launches_{tag}.csv")
    if os.path.exists(lst):
        lines = [l for l in open(lst) if not l.startswith("==")]
        rows = list(csv.DictReader(io.StringIO("".join(lines))))
        agg = {}
        for r in rows:
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            v = v / 1e6 if r["Metric Unit"] in ("ns", "nsecond") else (v / 1e3 if r["Metric Unit"] in ("us", "usecond") else v)
            a = agg.setdefault(r["Kernel Name"], [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(a[1] for a in agg.values())
        with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_launches_summary.csv"), "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow(["kernel", "launches", "total_ms", "share"])
            for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                w.writerow([k, a[0], round(a[1], 3), round(a[1] / tot, 4)])
        with gzip.open(os.path.join(ROOT, "profiles", f"{tag}_ncu_launches.csv.gz"), "wt") as fh:
            fh.write("".join(lines))
        print("launch list:", len(rows), "rows,", round(tot, 2), "ms under ncu")


if __name__ == "__main__":
    main()
