"""Build the checker: oracle/_build/libnms_oracle.so (gcc) and, where /root/reference exists,
oracle/_ref/libref_nms_kernel.so = the reference's own lib/nms/src/cuda/nms_kernel.cu compiled
UNCHANGED for sm_100a from where it lies (outputs only under oracle/_ref/, git-ignored, shipped to
the GPU box by gpurun).  nms.c / nms_cuda.c are unbuildable (TH/THC API removed) -- see DESIGN.md.
"""
import os
import subprocess

from . import nms_oracle, refshim

HERE = os.path.dirname(os.path.abspath(__file__))


def build_ref():
    src = os.path.join(refshim.REF_ROOT, "lib", "nms", "src", "cuda", "nms_kernel.cu")
    out = os.path.join(HERE, "_ref", "libref_nms_kernel.so")
    if not os.path.exists(src):
        return out if os.path.exists(out) else None
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared",
                               "-Xcompiler", "-fPIC", "-I", os.path.dirname(src), "-o", out, src, "-lcudart"])
    return out


def build():
    return nms_oracle.build(), build_ref()


if __name__ == "__main__":
    print(build())
