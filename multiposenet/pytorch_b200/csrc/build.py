"""Build libmpn_b200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["mpn_elementwise.cu", "mpn_conv_simt.cu", "mpn_conv_tc.cu", "mpn_detect.cu", "mpn_train.cu", "mpn_wgrad_tc.cu", "mpn_peaks.cu", "mpn_prn.cu", "mpn_resize.cu", "mpn_focal.cu"]
OUT = os.path.join(HERE, "libmpn_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v" if os.environ.get("MPN_PTXAS_V") else "-O3"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, s) for s in SOURCES] + [os.path.join(HERE, "mpn_common.cuh"), os.path.join(HERE, "mpn_tc_ptx.cuh"),
                                                        os.path.join(HERE, "..", "..", "..", "include", "mpn_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, trace=False):
    """trace=True: the instrumented debug build (-DMPN_CONV_TRACE: clock64 stamps inside conv_tc_kernel) -> libmpn_b200_trace.so,
    loaded through MPN_B200_LIB by scripts/exp/trace_conv.py; never the product library."""
    out_path = OUT.replace(".so", "_trace.so") if trace else OUT
    if not trace and not force and not needs_build():
        return OUT
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(HERE, s.replace(".cu", ".trace.o" if trace else ".o"))
        objs.append(o)
        cmd = [NVCC] + FLAGS + (["-DMPN_CONV_TRACE"] if trace else []) + ["-c", os.path.join(HERE, s), "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
        if verbose and out.strip():
            print(out)
    cmd = [NVCC, "-shared", "-o", out_path] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    return out_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, trace="--trace" in sys.argv))
