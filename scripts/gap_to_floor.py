"""Per-shape gap of one eager step to its floors, from the CUDA-event table of scripts/profile_layers.py.

    python scripts/gap_to_floor.py profiles/r02u_layers_f16f8_b32.txt [--batch 32] [--tflops 1354.1] [--gbs 6556.2] > profiles/r02u_gap_to_floor_f16f8.txt

tensor floor = 2 passes x 2 MAC / sustained bf16 peak (the f16f8 operand split costs two bf16-equivalent passes per MAC);
HBM floor   = compulsory bytes / measured HBM peak: input + output (+ shortcut / upsample source) at 3 B per element (fp16 + e5m2
              residual: the planes every consumer needs), fp32 head outputs 4 B, filters ignored.  No profiler involved.
"""
import argparse
import re


def parse(name):
    m = re.match(r"(multi\[(?P<lv>[^\]]+)\]|(?P<h>\d+)x(?P<w>\d+)) c(?P<ci>\d+)->(?P<co>\d+) k(?P<k>\d+) s(?P<s>\d+)(?P<rest>.*)", name)
    if not m:
        return None
    lv = [tuple(int(v) for v in t.split("x")) for t in m.group("lv").split(",")] if m.group("lv") else [(int(m.group("h")), int(m.group("w")))]
    return lv, int(m.group("ci")), int(m.group("co")), int(m.group("k")), int(m.group("s")), m.group("rest")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("table")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--tflops", type=float, default=1354.1)
    ap.add_argument("--gbs", type=float, default=6556.2)
    a = ap.parse_args()
    rows = []
    for line in open(a.table):
        m = re.match(r"^(.*?)\s+(\d+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s*$", line.rstrip())
        if not m:
            continue
        p = parse(m.group(1).strip())
        if p is None:
            continue
        lv, ci, co, k, s, rest = p
        cnt, ms_each = int(m.group(2)), float(m.group(4))
        # the table names a launch by its INPUT map (a stride-s launch writes an s x smaller output)
        ipix = sum(h * w for h, w in lv) * a.batch
        opix = sum(((h + s - 1) // s) * ((w + s - 1) // s) for h, w in lv) * a.batch
        mac = opix * ci * co * k * k
        tfl = 2.0 * 2.0 * mac / (a.tflops * 1e12) * 1e3
        out_b = 4.0 if co <= 36 else 3.0
        rep = int(re.search(r"rep(\d)", rest).group(1)) if "rep" in rest else 1
        byt = ipix * ci * 3.0 + opix * rep * rep * co * out_b
        if "res" in rest:
            byt += opix * co * 3.0
        if "up" in rest:
            byt += opix / 4 * co * 3.0
        hfl = byt / (a.gbs * 1e9) * 1e3
        fl = max(tfl, hfl)
        rows.append((m.group(1).strip(), cnt, ms_each, tfl, hfl, fl, ms_each / fl, cnt * (ms_each - fl)))
    rows.sort(key=lambda r: -r[7])
    print("f16f8, batch %d, one eager step (CUDA events per launch, %s) against the two floors of each launch:" % (a.batch, a.table))
    print("  tensor floor = 2 passes x 2*MAC / %.1f TFLOP/s (sustained cuBLAS bf16, MEASURED_PEAKS.json); HBM floor = compulsory bytes / %.0f GB/s" % (a.tflops, a.gbs))
    print("  (input + output [+ shortcut / upsample source], 3 B per element: fp16 + e5m2 residual; fp32 heads 4 B; filters ignored)")
    print("  ratios below 1: the sustained cuBLAS figure is a long-run, power-capped number that short launches beat in bursts, and the")
    print("  stem (7x7x3 taps inside a 4x64 K window) and phase-class filters (9 classes, most with empty taps) contain structural zeros")
    print("%-52s %5s %9s %9s %9s %9s %8s %8s" % ("launch", "count", "ms each", "tensor fl", "HBM fl", "floor", "x floor", "gap ms"))
    tot = tot_fl = 0.0
    for n, c, ms, tfl, hfl, fl, x, gap in rows:
        print("%-52s %5d %9.4f %9.4f %9.4f %9.4f %8.2f %8.3f" % (n[:52], c, ms, tfl, hfl, fl, x, gap))
        tot += c * ms
        tot_fl += c * fl
    print("%-52s %5s %9.3f %29s %9.3f %8.2f %8.3f" % ("total (ms per step)", "", tot, "", tot_fl, tot / tot_fl, tot - tot_fl))


if __name__ == "__main__":
    main()
