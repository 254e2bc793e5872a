"""Build libmnb200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
LIB = os.path.join(HERE, "libmnb200.so")
SOURCES = ["elementwise.cu", "dwconv.cu", "dwconv_tile.cu", "stem.cu", "conv_simt.cu", "gemm_tc.cu", "pw_stream.cu", "dw_mma.cu", "pw_bwd_fused.cu", "c3_mma.cu", "dw_small.cu", "stem_mma.cu", "pw_proj_bwd.cu", "pw_wide_fwd.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


STAMP = LIB + ".srchash"


def _source_hash():
    import hashlib
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + \
        [os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "mnb200.h")]
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def needs_build():
    """Content-based (file mtimes do not survive the copy to the GPU box)."""
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _source_hash()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    os.makedirs(os.path.join(CSRC, "build"), exist_ok=True)
    for s in SOURCES:
        o = os.path.join(CSRC, "build", s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc] + [f for f in FLAGS if not f.startswith("--use_fast_math")] + \
              ["-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcuda"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(STAMP, "w") as f:
        f.write(_source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
