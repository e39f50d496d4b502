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


def test_argument_errors_are_reported_not_crashed():
    """Every entry point checks its arguments before touching memory: null pointers with n > 0, negative sizes, bad enums
    and inconsistent search configurations come back as IFX_E_ARG (-1) -- through the CPU build of the same sources."""
    import ctypes as C
    import numpy as np
    from hostsim import hostsim_backend
    from inflexion_b200 import _capi
    dll = hostsim_backend().dll
    buf = np.zeros(4096, dtype=np.uint8)
    p = buf.ctypes.data
    E = _capi.IFX_E_ARG if hasattr(_capi, "IFX_E_ARG") else -1
    assert dll.ifx_rules_mask_step(None, None, 5, None, None, None, None) == E
    assert dll.ifx_rules_mask_step(p, None, -1, p, None, None, None) == E
    assert dll.ifx_rules_mask_step_host(None, None, 3, None, None, None) == E
    assert dll.ifx_rules_planes(None, None, 4, 0, None, None) == E
    assert dll.ifx_rules_planes(p, None, 4, 7, p, None) == E                      # dtype code
    assert dll.ifx_rules_symmetries(p, 1, 0, 4, p, None) == E                     # channels
    assert dll.ifx_rules_symmetries(p, 1, 4, 3, p, None) == E                     # element size
    assert dll.ifx_rules_transform(p, None, 1, 4, 4, p, None) == E                # no symmetry codes
    assert dll.ifx_rules_random_actions(None, 3, 1, 0, None, None) == E
    assert dll.ifx_rules_greedy_actions(None, 3, None, None) == E
    base = dict(n_games=2, nodes_per_game=256, hash_per_game=512, max_depth=352, num_sims=10, cpuct=1.0, seed=1,
                policy_is_logits=0, terminal_sims_per_wave=1, dirichlet_alpha=0.0, dirichlet_eps=0.0, leaves_per_wave=1)
    cfg, lay = _capi.MctsConfig(**base), _capi.MctsLayout()
    assert dll.ifx_mcts_get_layout(C.byref(cfg), C.byref(lay)) == 0 and dll.ifx_mcts_workspace_bytes(C.byref(cfg)) > 0
    bad = ({"n_games": 0}, {"n_games": -3}, {"nodes_per_game": 0}, {"nodes_per_game": 16384}, {"hash_per_game": 100},
           {"max_depth": 0}, {"max_depth": 10000}, {"num_sims": -1}, {"num_sims": 1 << 17}, {"leaves_per_wave": 17},
           {"leaves_per_wave": -2}, {"dirichlet_eps": 2.0}, {"dirichlet_eps": 0.25, "dirichlet_alpha": 0.0},
           {"leaves_per_wave": 4, "dirichlet_eps": 0.25, "dirichlet_alpha": 0.3})
    for kw in bad:
        cfg = _capi.MctsConfig(**{**base, **kw})
        h = C.c_void_p()
        assert dll.ifx_mcts_get_layout(C.byref(cfg), C.byref(lay)) == E, kw
        assert dll.ifx_mcts_workspace_bytes(C.byref(cfg)) == 0, kw
        assert dll.ifx_mcts_create(C.byref(cfg), p, 4096, C.byref(h)) == E and not h.value, kw
    cfg, h = _capi.MctsConfig(**base), C.c_void_p()
    assert dll.ifx_mcts_create(C.byref(cfg), p, 4096, C.byref(h)) == E            # workspace too small
    assert dll.ifx_mcts_wave(None, None, 0, None, None, None) == E                # null handle
    assert dll.ifx_mcts_status(None, None, None) == E
    assert dll.ifx_mcts_set_params(None, 5, 1.0) == E
    assert dll.ifx_mcts_root_stats(None, None, None, None, None, None, None) == E
    dll.ifx_mcts_destroy(None)                                                    # a no-op
