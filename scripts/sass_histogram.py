"""Instruction histogram of the built objects (cuobjdump -sass): the Blackwell-specific mnemonics that prove the kernels are
tcgen05 / TMA code (B200_PROFILING.md), per object file.  Runs without a GPU:
    python scripts/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import glob
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "multiposenet", "pytorch_b200", "csrc")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACMDFLUSH", "SYNCS", "UCGABAR", "LDGSTS",
        "HMMA", "FFMA", "F2FP", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "BAR", "SHFL"]


def main():
    objs = sorted(glob.glob(os.path.join(CSRC, "*.o")))
    if not objs:
        sys.exit("build first: python -m multiposenet.pytorch_b200.csrc.build")
    print("cuobjdump -sass instruction histogram (sm_100a objects of libmpn_b200.so); columns = mnemonic families (prefix match)")
    for o in objs:
        sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
        fn = None
        per_fn = collections.OrderedDict()
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                fn = m.group(1)
                per_fn[fn] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
            if m and fn:
                per_fn[fn][m.group(1)] += 1
        tot = collections.Counter()
        for c in per_fn.values():
            tot.update(c)
        fam = lambda c: {k: sum(v for op, v in c.items() if op.startswith(k)) for k in KEYS}
        t = fam(tot)
        print("\n== %s: %d kernels, %d instructions" % (os.path.basename(o), len(per_fn), sum(tot.values())))
        print("   " + "  ".join("%s=%d" % (k, v) for k, v in t.items() if v))
        sub = {op: v for op, v in tot.items() if op.startswith(("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM"))}
        if sub:
            print("   variants: " + "  ".join("%s=%d" % kv for kv in sorted(sub.items())))
        big = sorted(per_fn.items(), key=lambda kv: -sum(kv[1].values()))[:4]
        for name, c in big:
            f = fam(c)
            short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:150]
            print("   - %s\n       %d instr: %s" % (short, sum(c.values()), "  ".join("%s=%d" % (k, v) for k, v in f.items() if v)))


if __name__ == "__main__":
    main()
