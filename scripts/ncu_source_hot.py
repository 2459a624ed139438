"""Hot source lines of one kernel from an ncu report: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::K`
piped to a file, then `python scripts/ncu_source_hot.py FILE [top]`.  Prints per source line: samples, share, top stall reasons."""
import csv
import sys


def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    fname, hdr, agg = None, None, {}
    cur = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ix = {}
            for i, h in enumerate(hdr):
                ix.setdefault(h, i)
            stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0] != "":
            cur = (fname, int(r[0]), r[1].strip())
            continue
        if cur is None:
            continue
        a = agg.setdefault(cur, {"n": 0, "inst": 0, "st": {}})
        a["n"] += num(r[ix["# Samples"]])
        a["inst"] += num(r[ix["Instructions Executed"]])
        for s in stalls:
            v = num(r[ix[s]])
            if v:
                a["st"][s[6:]] = a["st"].get(s[6:], 0) + v
    tot = sum(a["n"] for a in agg.values())
    print("total samples", tot)
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:top]:
        st = sorted(a["st"].items(), key=lambda kv: -kv[1])[:3]
        print("%5.1f%% %7d inst %8d  %s:%d  %s  %s" % (100.0 * a["n"] / max(tot, 1), a["n"], a["inst"], k[0], k[1], k[2][:90], st))


if __name__ == "__main__":
    main()
