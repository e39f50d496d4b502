"""The C ABI: every symbol include/ifx_b200.h declares is bound by the package and exported by the built sm_100a
library (no compute calls here: this runs without a GPU), and the product refuses to run without CUDA."""
import ctypes
import os
import re
import shutil

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ifx_b200.h")).read()
    return sorted(set(re.findall(r"IFX_API\s+[\w\s\*]+?\b(ifx_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    from inflexion_b200 import _capi
    syms = declared_symbols()
    assert len(syms) >= 30
    assert sorted(_capi.PROTOTYPES) == syms


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="no nvcc")
def test_built_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()                                            # nvcc -gencode arch=compute_100a,code=sm_100a ...
    dll = ctypes.CDLL(entry.LIB)
    for name in declared_symbols():
        assert hasattr(dll, name), name
    dll.ifx_version.restype = ctypes.c_int
    dll.ifx_is_hostsim.restype = ctypes.c_int
    assert dll.ifx_version() >= 100 and dll.ifx_is_hostsim() == 0


def test_state_struct_is_64_bytes():
    from inflexion_b200.state import STATE_DTYPE
    assert STATE_DTYPE.itemsize == 64
    assert [STATE_DTYPE.fields[k][1] for k in ("red", "p2", "turn", "player", "outcome", "total_power", "tag")] == \
        [0, 32, 40, 44, 45, 46, 48]


def test_product_has_no_cpu_fallback(monkeypatch):
    import torch
    import inflexion_b200.backend as backend_mod
    from inflexion_b200._capi import IfxError
    from inflexion_b200.inflexion.InflexionGame import InflexionGame
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    monkeypatch.setattr(backend_mod, "_DEFAULT", None)
    with pytest.raises(IfxError):
        InflexionGame(7).valid_actions_mask()                # needs libifx_b200.so AND a CUDA device


def test_header_is_plain_c_and_the_c_example_compiles(tmp_path):
    """include/ifx_b200.h must be usable from C (no C++-isms, no CUDA or torch headers): examples/c_abi_demo.c, a C99
    host program that dlopen()s the library, has to compile with gcc alone."""
    import subprocess
    out = tmp_path / "c_abi_demo"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", str(out), "-ldl"]
    subprocess.run(cmd, check=True, capture_output=True)
    assert out.exists()
