"""GPU-less logic check of the kernels: compiles alpha-zero-general-inflexion_b200/csrc/*.cu as plain C++
(-DIFX_HOSTSIM: every warp emulated as 32 sequential lanes) into tests/_build/ and returns a Backend that runs
it on CPU tensors.  TEST INFRASTRUCTURE ONLY -- lets the descent / backup / hashing / rounding logic be
compared with the oracle in the build container, which has no GPU.  The product never loads this library."""
import glob
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "alpha-zero-general-inflexion_b200")
OUT = os.path.join(ROOT, "tests", "_build", "libifx_hostsim.so")


def build_hostsim() -> str:
    srcs = sorted(glob.glob(os.path.join(PKG, "csrc", "*.cu")))
    deps = srcs + glob.glob(os.path.join(PKG, "csrc", "*.cuh")) + [os.path.join(ROOT, "include", "ifx_b200.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-x", "c++", "-DIFX_HOSTSIM", "-O2", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas",
           "-shared", "-fPIC", "-fvisibility=hidden", "-I", os.path.join(ROOT, "include"), "-o", OUT] + srcs
    subprocess.run(cmd, check=True)
    return OUT


def hostsim_backend():
    import inflexion_b200  # noqa: F401
    from inflexion_b200 import _capi
    from inflexion_b200.backend import Backend
    dll = _capi.bind(build_hostsim())
    assert dll.ifx_is_hostsim() == 1
    return Backend(dll, "cpu")
