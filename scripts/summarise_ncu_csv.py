"""Summarise an `ncu --csv --metrics ...` log (one row per launch and metric) into per-kernel totals and a JSON record.

    python scripts/summarise_ncu_csv.py gpurun_out/conv_traffic.csv --kernel conv_tc_kernel --out profiles/r01p_conv_traffic.json
bench.py reads the newest profiles/*_conv_traffic.json for `roofline.traffic` (DRAM bytes per conv launch, averaged
over the launches of one forward step)."""
import argparse
import csv
import gzip
import io
import json
import re


def read_rows(path):
    raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
    start = raw.find('"ID"')
    return list(csv.DictReader(io.StringIO(raw[start:])))


def to_float(v):
    return float(v.replace(",", "")) if v not in ("", "n/a", None) else float("nan")


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3,
        "nsecond": 1e-9, "second": 1.0, "%": 1.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--kernel", default="conv_tc_kernel")
    ap.add_argument("--out", default=None)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    launches = {}
    for r in read_rows(a.csv):
        name = r.get("Kernel Name", "")
        if a.kernel not in name:
            continue
        d = launches.setdefault(r["ID"], {"kernel": re.sub(r"\(.*", "", name)})
        d[r["Metric Name"]] = to_float(r["Metric Value"]) * UNIT.get(r.get("Metric Unit", ""), 1.0)
    n = len(launches)
    rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in launches.values())
    wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in launches.values())
    t = sum(d.get("gpu__time_duration.sum", 0.0) for d in launches.values())
    key = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    tens = sum(d.get(key, 0.0) * d.get("gpu__time_duration.sum", 0.0) for d in launches.values()) / t if t else None
    rec = {"kernel": a.kernel, "launches": n, "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes_per_launch": (rd + wr) / max(n, 1),
           "time_s_under_ncu": t, "dram_gbs_under_ncu": (rd + wr) / t / 1e9 if t else None, "tensor_pipe_active_pct_time_weighted": tens,
           "note": a.note}
    print(json.dumps(rec, indent=1))
    if a.out:
        json.dump(rec, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
