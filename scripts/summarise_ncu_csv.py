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
    ap.add_argument("--names", default=None, help="layer shape per launch, in launch order (scripts/ncu_conv_step.py --names-out)")
    ap.add_argument("--table", default=None, help="with --names: per-shape table (count, us, DRAM GB/s, %% of HBM peak, tensor-pipe %%)")
    ap.add_argument("--hbm-gbs", type=float, default=6556.2, help="measured HBM peak (MEASURED_PEAKS.json)")
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
    if a.names:
        names = [l.strip() for l in open(a.names) if l.strip()]
        ids = sorted(launches, key=lambda k: int(k))
        if len(names) != len(ids):
            raise SystemExit("--names has %d rows, the CSV %d launches of %s" % (len(names), len(ids), a.kernel))
        agg = {}
        for nm, i in zip(names, ids):
            d = launches[i]
            g = agg.setdefault(nm, [0, 0.0, 0.0, 0.0])
            tt = d.get("gpu__time_duration.sum", 0.0)
            g[0] += 1; g[1] += tt; g[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0); g[3] += d.get(key, 0.0) * tt
        lines = ["%-38s %5s %10s %9s %10s %9s %9s" % ("launch (under ncu: serialised, cold L2)", "count", "us total", "us each", "DRAM GB/s", "of HBM", "tensor %")]
        for nm, g in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            gbs = g[2] / g[1] / 1e9 if g[1] else 0.0
            lines.append("%-38s %5d %10.1f %9.1f %10.0f %8.1f%% %8.1f%%" % (nm, g[0], g[1] * 1e6, g[1] * 1e6 / g[0], gbs, 100.0 * gbs / a.hbm_gbs, g[3] / g[1] if g[1] else 0.0))
        txt = "\n".join(lines)
        print(txt)
        if a.table:
            open(a.table, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
