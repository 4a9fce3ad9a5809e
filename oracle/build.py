"""TEST INFRASTRUCTURE ONLY.  Build recipe for the checker libraries:

  oracle/liboracle.so            <- oracle/oracle.c (our plain-C restatement)
  oracle/_ref/libdnn_openblas.so <- /root/reference/cs492-projects/proj3/dnn_openblas.c, unmodified
  oracle/_ref/libdnn_avx.so      <- /root/reference/cs492-projects/proj3/dnn_avx.c, unmodified

The two reference libraries are compiled from the sources where they lie (never copied),
with the flags of proj3/Makefile:4,7 plus oracle/shim/cblas.h, which maps the three CBLAS
calls onto the OpenBLAS bundled with scipy.  They are only (re)built when /root/reference
exists, i.e. in the build container; the GPU box uses the prebuilt files that travel with
the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/cs492-projects/proj3"
REF_OUT = os.path.join(HERE, "_ref")


def _run(cmd):
    subprocess.run(cmd, check=True)


def _newer(target, *sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def scipy_openblas():
    import scipy
    pat = os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas-*.so")
    libs = glob.glob(pat)
    if not libs:
        raise RuntimeError("scipy-bundled OpenBLAS not found: " + pat)
    return os.path.realpath(libs[0])


def build_oracle(verbose=False):
    src = os.path.join(HERE, "oracle.c")
    out = os.path.join(HERE, "liboracle.so")
    if _newer(out, src):
        # no -march=native: the .so travels to a box whose CPU may differ
        _run(["gcc", "-shared", "-fPIC", "-O3", "-mavx2", "-mfma", "-fopenmp", "-o", out, src, "-lm"])
        if verbose:
            print("built", out)
    return out


def build_ref(verbose=False):
    """Returns the list of reference libraries present after the call (may be empty)."""
    os.makedirs(REF_OUT, exist_ok=True)
    ob_so = os.path.join(REF_OUT, "libdnn_openblas.so")
    avx_so = os.path.join(REF_OUT, "libdnn_avx.so")
    if os.path.isdir(REF_SRC):
        shim = os.path.join(HERE, "shim")
        ob = scipy_openblas()
        src = os.path.join(REF_SRC, "dnn_openblas.c")
        if _newer(ob_so, src, os.path.join(shim, "cblas.h")):
            _run(["gcc", "-shared", "-fPIC", "-I", shim, "-o", ob_so, src,
                  "-L", os.path.dirname(ob), "-l:" + os.path.basename(ob),
                  "-Wl,-rpath," + os.path.dirname(ob), "-lm"])
            if verbose:
                print("built", ob_so)
        src = os.path.join(REF_SRC, "dnn_avx.c")
        if _newer(avx_so, src, os.path.join(shim, "cblas.h")):
            _run(["gcc", "-shared", "-fPIC", "-O3", "-w", "-I", shim, "-o", avx_so, src, "-mavx2", "-lpthread"])
            if verbose:
                print("built", avx_so)
    return [p for p in (ob_so, avx_so) if os.path.exists(p)]


if __name__ == "__main__":
    build_oracle(verbose=True)
    print("reference libs:", build_ref(verbose=True))
