"""The C-ABI library loads, exports exactly what include/sda_b200.h declares, and has no CPU
fallback.  No compute calls (CPU suite)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import sda_b200
from sda_b200 import _lib
from sda_b200 import LinearMaskingScheme as LMS
from sda_b200 import LinearSecretSharingScheme as LSS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sda_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sda_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = sda_b200.load()
    names = declared_symbols()
    assert len(names) >= 35
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/sda_b200.h but not exported"
    # and the ctypes prototype table covers the header, nothing more, nothing less
    assert sorted(_lib.PROTOTYPES) == names
    out = subprocess.run(["nm", "-D", "--defined-only", sda_b200.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(re.findall(r" T (sda_[a-z0-9_]+)", out))
    assert exported == names


def test_abi_version():
    assert sda_b200.load().sda_abi_version() == 1


def test_struct_layout_matches_header():
    assert C.sizeof(_lib.sda_sharing_scheme) == 56
    assert C.sizeof(_lib.sda_masking_scheme) == 32


def test_scheme_sizes():
    """protocol/src/crypto.rs:117-155 through the ABI (pure host functions)"""
    a = LSS.Additive(3, 433)
    assert (a.input_size(), a.output_size(), a.privacy_threshold(), a.reconstruction_threshold()) == (1, 3, 2, 3)
    p = LSS.PackedShamir(3, 8, 4, 433, 354, 150)
    assert (p.input_size(), p.output_size(), p.privacy_threshold(), p.reconstruction_threshold()) == (3, 8, 4, 7)
    assert p.batches(4) == 2 and p.batches(0) == 0 and a.batches(10) == 10
    assert LMS.None_().mask_len(10) == 0 and LMS.Full(433).mask_len(10) == 10
    assert LMS.ChaCha(433, 10, 128).mask_len(10) == 4 and LMS.ChaCha(433, 10, 40).mask_len(10) == 2
    assert not LMS.None_().has_mask() and LMS.Full(433).has_mask()


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product refuses to run instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sda_b200.SdaClientError) as e:
        sda_b200.CryptoModule(0)
    assert e.value.code == _lib.SDA_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under sda_b200/ may reference it."""
    pkg = os.path.join(ROOT, "sda_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "sda_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
    out = subprocess.run(["ldd", sda_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
