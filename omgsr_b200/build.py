"""Builds libomgsr_b200.so (hand-written sm_100a CUDA behind the C ABI in include/omgsr_b200.h) in-tree
with a plain nvcc invocation — no torch dependency in the library, no JIT cache."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libomgsr_b200.so")
SOURCES = ["capi.cu", "gemm_conv_tc.cu", "conv_halo_tc.cu", "attn_tc.cu", "glue.cu", "colorfix.cu", "resize.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math", "-Xptxas", "-v",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    # every header the sources include: editing epilogue.cuh (or any other .cuh / .h) must trigger a rebuild
    inc = os.path.join(HERE, "..", "include")
    deps = srcs + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + \
        sorted(os.path.join(inc, f) for f in os.listdir(inc) if f.endswith(".h"))
    stamp = os.path.join(LIBDIR, "build.sha256")
    dig = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    objs = []
    logs = []
    procs = []
    for s in srcs:
        o = os.path.join(LIBDIR, os.path.basename(s).replace(".cu", ".o"))
        objs.append(o)
        procs.append((s, subprocess.Popen([_nvcc()] + NVCC_FLAGS + ["-c", s, "-o", o], stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        logs.append(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {s}")
    link = subprocess.run([_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart"], capture_output=True, text=True)
    if link.returncode != 0:
        sys.stderr.write(link.stdout + link.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    with open(stamp, "w") as f:
        f.write(dig)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
