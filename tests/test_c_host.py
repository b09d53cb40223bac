"""The drop-in boundary from a host that is not Python: examples/host_render.c is plain C99 against include/gmpi_mpi_render.h
and the shared library only (no toolkit, no torch).  CPU: the header is valid pedantic C, the program links, and without a GPU
the library fails loudly with a message (no CPU fallback).  GPU: the program renders from host memory and verifies closed-form
answers of the compositing rule itself (exit code 0)."""
import os
import shutil
import subprocess

import pytest

import ml_gmpi_b200 as g
from conftest import ROOT


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    g.build_library()
    out = str(tmp_path_factory.mktemp("c_host") / "host_render")
    lib_dir = os.path.join(ROOT, "ml_gmpi_b200")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "host_render.c"), "-o", out, "-L", lib_dir, "-lgmpi_mpi_render", f"-Wl,-rpath,{lib_dir}", "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return out


def test_c_host_links_and_fails_loudly_without_a_gpu(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 3, (res.stdout, res.stderr)
    assert "gmpi ABI version 2" in res.stdout and "gmpi_mpi_render_fwd_host failed" in res.stderr


@pytest.mark.gpu
def test_c_host_renders_and_verifies(exe):
    res = subprocess.run([exe, "0"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, (res.stdout, res.stderr)
    assert res.stdout.strip().endswith("OK") and res.stdout.count("max abs error") == 8
