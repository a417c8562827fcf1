"""Builds libhl_b200.so (CUDA, sm_100a) in-tree with nvcc.  Called by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libhl_b200.so")
SOURCES = ["hlb_api.cu", "hlb_batch.cu", "hlb_slice.cu"]
# -dlcm=cg: plain global loads are served from L2 (macroblock state / reconstruction written by other CTAs of the slice kernel);
# read-only planes use __ldg explicitly.  --fmad=false: the RD cost is a double sum that must round like the reference's.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=false",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-Xptxas", "-v,-dlcm=cg"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "hlb200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = ["-DHLB_SLICE_MIN_CTAS=" + os.environ["HLB_SLICE_MIN_CTAS"]] if os.environ.get("HLB_SLICE_MIN_CTAS") else []
    if os.environ.get("HLB_WORKERS"):
        extra.append("-DHLB_WORKERS=" + os.environ["HLB_WORKERS"])
    extra += os.environ.get("HLB_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out or OUT]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed, see %s" % log)
    return OUT


if __name__ == "__main__":
    o = [a for a in sys.argv[1:] if a.endswith(".so")]
    build(force=True, verbose="-v" in sys.argv, out=o[0] if o else None)
    print(o[0] if o else OUT)
