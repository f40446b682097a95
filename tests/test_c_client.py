"""The boundary is a plain C ABI: a C program (tests/c/abi_smoke.c, gcc, no C++) includes
include/hawkscan.h, links libhawkscan.so and runs. CPU: host-only entry points. GPU: one
KAT1 search end to end with the known answers of SURVEY.md Appendix A."""

import os
import shutil
import subprocess

import pytest

from crispr_hawk_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    build.build_cuda()
    exe = str(tmp_path / "abi_smoke")
    cc = shutil.which("gcc") or "gcc"
    libdir = os.path.dirname(_cabi.LIB_PATH)
    subprocess.run(
        [cc, "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
         "-o", exe, "-L", libdir, "-lhawkscan", f"-Wl,-rpath,{libdir}"],
        check=True,
    )  # fmt: skip
    return exe


def test_c_client_host_entry_points(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "layout ok" in out.stdout


@pytest.mark.gpu
def test_c_client_search(tmp_path):
    out = subprocess.run([_build(tmp_path), "gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "guides 14 hits 3/11 window 43 scanned 77" in out.stdout
    assert "annotate 1 stream 1 capacity 1 traffic 1" in out.stdout
