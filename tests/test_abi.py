"""CPU: libsvx.so builds, loads and exports exactly the symbols include/svx.h declares.  No
compute call is made here (there is no GPU in the build container)."""
import os
import re
import subprocess

import pytest

from svision_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "svx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svx_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(lib_path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path]).decode()
    exported = set(re.findall(r" T (svx_[a-z0-9_]+)", out))
    assert exported == set(_lib.SYMBOLS)


def test_library_loads_and_reports_version(lib_path):
    lib = _lib.load()
    assert b"sm_100a" in lib.svx_version()
    assert lib.svx_launch_count() == 0


def test_library_is_blackwell_native(lib_path):
    sass = subprocess.check_output(["cuobjdump", "-sass", lib_path]).decode()
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):      # tcgen05.mma, TMA load, tcgen05.ld
        assert mnemonic in sass, mnemonic
    assert "HMMA.16" not in sass                          # no legacy mma.sync path


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from svision_b200 import classifier
    with pytest.raises(_lib.SvxError):
        classifier.Classifier(None)
