"""In-tree build of the native pieces (no JIT cache: the .so files travel with the repo).

* ``libhawkscan.so``   -- CUDA kernels + C-ABI, nvcc, sm_100a only
* ``libhawkcheck.so``  -- the same __host__ __device__ core compiled for the CPU
  (g++), used by the ``-m "not gpu"`` tests to exercise the kernels' bit logic
  without a device. Not a product path: it exports no search entry point.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libhawkscan.so")
CHECK_LIB = os.path.join(PKG, "libhawkcheck.so")

CUDA_SOURCES = ["scan_kernels.cu", "scan2_kernels.cu", "fused_kernels.cu", "post_kernels.cu", "resolve_kernels.cu", "table_kernels.cu", "synth_kernels.cu", "annot_kernels.cu", "api.cu", "stream_api.cu", "annot_api.cu", "merge_api.cu", "collapse_api.cu", "edits_kernels.cu", "cfdon_api.cu", "featurize_api.cu"]
HEADERS = ["hawk_core.h", "hawk_kernels.h", "hawk_post.h", "hawk_host.h", "../../include/hawkscan.h"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


class _BuildLock:
    """Serialises concurrent builds (torchrun starts one process per GPU)."""

    def __enter__(self):
        import fcntl

        self.f = open(os.path.join(PKG, ".build.lock"), "w")
        fcntl.flock(self.f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *a):
        import fcntl

        fcntl.flock(self.f, fcntl.LOCK_UN)
        self.f.close()


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    with _BuildLock():
        return _build_cuda(force, verbose)


def build_hostcheck(force: bool = False) -> str:
    with _BuildLock():
        return _build_hostcheck(force)


def build_variant(out: str, defines) -> str:
    """Tuning aid: the same sources with -D overrides into another .so (HAWKSCAN_LIB selects it)."""
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC,-O2", "--expt-relaxed-constexpr", "-o", out] + [f"-D{d}" for d in defines] + srcs
    subprocess.run(cmd, check=True, cwd=CSRC)
    return out


def _build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = srcs + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    if not force and not _newer(LIB, deps):
        return LIB
    cmd = [
        nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
        "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unknown-pragmas", "--expt-relaxed-constexpr",
        "-o", LIB,
    ] + srcs  # fmt: skip
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB


def _build_hostcheck(force: bool = False) -> str:
    src = os.path.join(CSRC, "hostcheck.cpp")
    deps = [src, os.path.join(CSRC, "hawk_core.h"), os.path.normpath(os.path.join(CSRC, "../../include/hawkscan.h"))]
    if not force and not _newer(CHECK_LIB, deps):
        return CHECK_LIB
    cxx = shutil.which("g++") or "g++"
    subprocess.run(
        [cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", CHECK_LIB, src],
        check=True, cwd=CSRC,
    )  # fmt: skip
    return CHECK_LIB


if __name__ == "__main__":
    build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_hostcheck(force="--force" in sys.argv)
    print(LIB)
    print(CHECK_LIB)
