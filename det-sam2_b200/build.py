"""Builds libdetsam2.so (sm_100a only) in-tree with nvcc.

The shared object lands in ``det-sam2_b200/lib/libdetsam2.so`` so that it travels with the source
snapshot to the GPU box (no JIT cache).  Re-builds only when a source is newer than the library.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libdetsam2.so")

SOURCES = [
    "runtime.cu",
    "gemm_tc.cu",
    "flash_tc.cu",
    "mha.cu",
    "win_attn_tc.cu",
    "elementwise.cu",
    "conv.cu",
    "decoder.cu",
    "post.cu",
    "prompt.cu",
    "bank.cu",
    "ingest.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libdetsam2.so cannot be built")
    return nvcc


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "detsam2.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            raise RuntimeError(f"missing source {path}")
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(os.path.join(CSRC, f))
                                                for f in os.listdir(CSRC))
                and os.path.getmtime(obj) > os.path.getmtime(os.path.join(HERE, "..", "include", "detsam2.h"))):
            continue
        cmd = [nvcc] + NVCC_FLAGS + ["-c", path, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"---- nvcc failed on {src} ----\n{out}\n")
        elif verbose:
            sys.stderr.write(f"---- {src} ----\n{out}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
