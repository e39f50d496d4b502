"""ctypes binding of the C ABI declared in include/ifx_b200.h.

``lib()`` returns the product library (libifx_b200.so, nvcc sm_100a) and raises if it is missing -- there is
no CPU fallback.  ``bind(path)`` attaches the same prototypes to any build of the ABI (the tests use it for
the GPU-less logic-check build).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libifx_b200.so")

c_void_p, c_int, c_int32, c_int64, c_size_t, c_float, c_uint64 = (
    C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_size_t, C.c_float, C.c_uint64)


class MctsConfig(C.Structure):
    _fields_ = [("n_games", c_int32), ("nodes_per_game", c_int32), ("hash_per_game", c_int32),
                ("max_depth", c_int32), ("num_sims", c_int32), ("cpuct", c_float), ("seed", c_uint64),
                ("policy_is_logits", c_int32), ("terminal_sims_per_wave", c_int32),
                ("dirichlet_alpha", c_float), ("dirichlet_eps", c_float), ("leaves_per_wave", c_int32)]


class MctsLayout(C.Structure):
    _fields_ = [(n, c_int64) for n in ("rows", "q64", "node_key", "node_head", "hash", "free_stack", "per_game",
                                        "path", "roots", "leaves", "sym", "total")]


# name -> (restype, argtypes); every symbol include/ifx_b200.h declares
PROTOTYPES = {
    "ifx_version": (c_int, []),
    "ifx_last_cuda_error": (C.c_char_p, []),
    "ifx_launch_count": (c_int64, []),
    "ifx_is_hostsim": (c_int, []),
    "ifx_rules_mask_step": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ifx_rules_planes": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "ifx_rules_symmetries": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "ifx_rules_transform": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "ifx_rules_random_actions": (c_int, [c_void_p, c_int64, c_uint64, C.c_uint32, c_void_p, c_void_p]),
    "ifx_rules_greedy_actions": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "ifx_rules_info": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "ifx_rules_mask_step_host": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "ifx_rules_planes_host": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "ifx_rules_symmetries_host": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p]),
    "ifx_rules_transform_host": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "ifx_rules_info_host": (c_int, [c_void_p, c_int64, c_void_p]),
    "ifx_mcts_get_layout": (c_int, [C.POINTER(MctsConfig), C.POINTER(MctsLayout)]),
    "ifx_mcts_workspace_bytes": (c_size_t, [C.POINTER(MctsConfig)]),
    "ifx_mcts_create": (c_int, [C.POINTER(MctsConfig), c_void_p, c_size_t, C.POINTER(c_void_p)]),
    "ifx_mcts_destroy": (None, [c_void_p]),
    "ifx_mcts_set_params": (c_int, [c_void_p, c_int32, c_float]),
    "ifx_mcts_wave_timing": (c_int, [c_void_p, c_int]),
    "ifx_mcts_wave_time": (c_int, [c_void_p, C.POINTER(C.c_double), C.POINTER(c_int32)]),
    "ifx_mcts_set_debug": (c_int, [c_void_p, c_void_p]),
    "ifx_mcts_reset_games": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "ifx_mcts_set_roots": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ifx_mcts_wave": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "ifx_mcts_wave_part": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "ifx_mcts_root_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ifx_mcts_sample_actions": (c_int, [c_void_p, c_int32, c_void_p, c_void_p]),
    "ifx_mcts_advance": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ifx_mcts_restart_finished": (c_int, [c_void_p, C.c_uint32, c_void_p, c_void_p]),
    "ifx_mcts_status": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ifx_mcts_get_roots": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ifx_mcts_counters": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ifx_mcts_live_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "ifx_gather_pack_rows": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "ifx_gather_scatter_rows": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "ifx_net_im2col3x3": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "ifx_stub_policy_value": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p]),
}

ERRORS = {-1: "IFX_E_ARG (bad argument)", -2: "IFX_E_CUDA", -3: "IFX_E_NODEVICE"}
FAULTS = {1: "node pool exhausted (raise nodes_per_game)", 2: "no valid action", 4: "search path deeper than max_depth",
          8: "hash table full", 16: "invalid action", 32: "visit count reached 2^17"}


class IfxError(RuntimeError):
    pass


def bind(path: str) -> C.CDLL:
    dll = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(dll, name)          # AttributeError here = the build does not export the declared ABI
        fn.restype, fn.argtypes = res, args
    return dll


_LIB = None


def lib() -> C.CDLL:
    """The CUDA library.  Raises when it has not been built: the product has no CPU path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise IfxError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  inflexion_b200 has no CPU fallback.")
        _LIB = bind(LIB_PATH)
        if _LIB.ifx_is_hostsim():
            raise IfxError("libifx_b200.so is a host-simulation build; the product requires the sm_100a build")
    return _LIB


def check(rc: int, dll=None) -> None:
    if rc != 0:
        detail = ""
        if rc == -2:
            detail = ": " + ((dll or lib()).ifx_last_cuda_error() or b"").decode()
        raise IfxError(f"libifx_b200 call failed: {ERRORS.get(rc, rc)}{detail}")


def fault_text(bits: int) -> str:
    return ", ".join(t for b, t in FAULTS.items() if bits & b) or "none"
