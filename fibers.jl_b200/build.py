"""Build libfibers_cuda.so in-tree with nvcc for sm_100a (no JIT cache, no CPU fallback).

    python fibers.jl_b200/build.py [--force]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfibers_cuda.so")
BUILD = os.path.join(HERE, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread"]
# per-file extra flags
EXTRA = {
    "dti.cu": ["-fmad=false"],          # eigen-solver follows the reference's unfused fp32 order
    "structens.cu": ["-fmad=false"],
    "stream.cu": ["-fmad=false"],        # propagation follows the reference's unfused fp32 order
    "setup.cpp": ["-Xcompiler", "-ffp-contract=off"],
    "mri_io.cpp": ["-Xcompiler", "-ffp-contract=off"],
    "recon_simt.cu": ["-diag-suppress", "128"],   # "loop is not reachable" in the plain-rows instantiation (early return)
}
SOURCES = ["api.cu", "host_pipeline.cu", "setup.cpp", "dti.cu", "recon_simt.cu", "recon_tc.cu", "rumba.cu", "structens.cu", "stream.cu", "mri_io.cpp"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libfibers_cuda cannot be built (there is no CPU fallback)")


def _digest(paths: list[str], extra: str = "") -> str:
    """sha256 over the CONTENT of the inputs and the command line (mtimes do not survive a snapshot / checkout)."""
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(target: str, stamp: str, digest: str) -> bool:
    if not os.path.exists(target) or not os.path.exists(stamp):
        return True
    with open(stamp) as f:
        return f.read().strip() != digest


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles what is out of date (by content hash) and links the library; prints what it did."""
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "fibers_cuda.h"))
    objs, procs, stamps = [], [], {}
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, src.rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        cmd = [nvcc, *ARCH, *COMMON, *EXTRA.get(src, []), "-c", sp, "-o", obj]
        digest = _digest([sp] + headers, " ".join(cmd[1:-4]))
        if force or _stale(obj, obj + ".sha256", digest):
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            stamps[src] = (obj + ".sha256", digest)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed.append(f"--- {src} ---\n{out}")
        else:
            with open(stamps[src][0], "w") as f:
                f.write(stamps[src][1])
            if verbose or out.strip():
                print(f"--- {src} ---\n{out}")
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(failed))
    link_digest = _digest(objs, "link")
    relinked = False
    if force or procs or _stale(LIB, LIB + ".sha256", link_digest):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart", "-lcuda", "-ldl", "-lz", "-Xcompiler", "-pthread"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
        with open(LIB + ".sha256", "w") as f:
            f.write(link_digest)
        relinked = True
    print(f"[fibers build] sm_100a: compiled {[s for s, _ in procs] or 'nothing (content hashes match)'}; "
          f"{'linked' if relinked else 'kept'} {os.path.relpath(LIB, os.path.join(HERE, '..'))}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
