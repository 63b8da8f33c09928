"""tests/c_abi_smoke.c: the reference's end-to-end vectors through include/sda_b200.h from plain C (no Python in
the data path).  The CPU suite builds it with gcc and checks that, without a device, it reports "no CUDA device"
(exit 77) instead of computing anything; the GPU suite expects every vector to be reproduced."""
import os
import subprocess

import pytest

import sda_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    exe = str(tmp_path / "c_abi_smoke")
    libdir = os.path.dirname(sda_b200.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe, "-L", libdir, "-lsda_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def test_c_smoke_builds_and_refuses_to_run_without_a_device(tmp_path):
    import torch
    exe = build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: see the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 77, r.stdout + r.stderr
    assert "no CUDA device" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_c_smoke_reproduces_the_reference_vectors(tmp_path):
    r = subprocess.run([build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count(" ok") == 6 and "MISMATCH" not in r.stdout
