"""ctypes binding of libg2048.so (include/g2048.h).  Thin by design: pointers and sizes in,
error codes out.  There is no CPU fallback -- a missing library or device raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("G2048_LIB_PATH") or os.path.join(_HERE, "libg2048.so")   # override: kernel experiments
CSRC = os.path.join(_HERE, "csrc")

STATS_LEN = 40
STATS_MAXSCORE = 23
MAX_BEAM_WIDTH = 32          # widths up to this take the warp-shuffle fast path
MAX_WIDE_BEAM_WIDTH = 128    # accepted maximum (33 and up: slower shared-memory path)


class G2048Error(RuntimeError):
    pass


def build_library(force: bool = False, verbose: bool = False) -> str:
    """nvcc-compiles csrc/ for sm_100a into libg2048.so next to this file (no GPU needed)."""
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    out = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if verbose:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise G2048Error("building libg2048.so failed:\n" + out.stdout + out.stderr)
    return LIB_PATH


_lib = None
_ready_devices = set()

_vp, _i64, _i32, _u32, _u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_uint64

_SIGNATURES = {
    "g2048_abi_version": ([], C.c_int),
    "g2048_init": ([C.c_int], C.c_int),
    "g2048_set_device": ([C.c_int], C.c_int),
    "g2048_last_error": ([], C.c_char_p),
    "g2048_overflow_count": ([C.POINTER(_u64), _vp], C.c_int),
    "g2048_launch_count": ([], _u64),
    "g2048_set_tuning": ([C.c_int, C.c_int], C.c_int),
    "g2048_pack": ([_vp, _vp, _i64, _vp], C.c_int),
    "g2048_unpack": ([_vp, _vp, _i64, _vp], C.c_int),
    "g2048_observe": ([_vp, _vp, _i64, _vp], C.c_int),
    "g2048_simulate_move": ([_vp] * 7 + [_i64, _vp], C.c_int),
    "g2048_evaluate_pattern": ([_vp, _vp, _i64, _vp], C.c_int),
    "g2048_host_simulate_move": ([_vp] * 8 + [_i64], C.c_int),
    "g2048_hybrid_expand": ([_vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _u64, _u32, _vp], C.c_int),
    "g2048_hybrid_expand_items": ([_vp] * 10 + [_i64, _u64, _u32, _vp], C.c_int),
    "g2048_host_hybrid_expand": ([_vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _u64, _u32], C.c_int),
    "g2048_ppo_features": ([_vp, _vp, _vp, _vp, _i64, _vp], C.c_int),
    "g2048_host_ppo_features": ([_vp, _vp, _vp, _vp, _i64], C.c_int),
    "g2048_novelty_set_init": ([_vp, _vp, _i64, _vp], C.c_int),
    "g2048_ppo_shape_rewards": ([_vp] * 6 + [_i64, _u32, _vp, _vp, _vp, _i64, _vp], C.c_int),
    "g2048_synthetic_boards": ([_vp, _i64, _u64, _u32, _vp], C.c_int),
    "g2048_env_reset": ([_vp, _vp, _vp, _vp, _i64, _u64, _u32, _vp], C.c_int),
    "g2048_env_reset_done": ([_vp] * 6 + [_i64, _u64, _u32, _vp], C.c_int),
    "g2048_env_step_autoreset": ([_vp] * 12 + [_i64, _u64, _u32, _vp], C.c_int),
    "g2048_env_step": ([_vp] * 12 + [_i64, _u64, _u32, _vp], C.c_int),
    "g2048_env_step_fused": ([_vp] * 16 + [_i64, _u64, _u32, _vp], C.c_int),
    "g2048_legal_masks": ([_vp, _vp, _vp, _i64, _vp], C.c_int),
    "g2048_env_rollout": ([_vp] * 6 + [_i64, _i32, _u32, _u64, _u32, _vp], C.c_int),
    "g2048_evaluate": ([_vp, _vp, _vp, _i64, _vp], C.c_int),
    "g2048_beam_search": ([_vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _u64, _u32, _vp], C.c_int),
    "g2048_play_games": ([_i64, _i32, _i32, _i32, _i32, _i32, _u64, _u32] + [_vp] * 8 + [_vp], C.c_int),
    "g2048_stats_reduce": ([_vp] * 7 + [_i64, _vp, _vp], C.c_int),
    "g2048_host_env_reset": ([_vp, _vp, _vp, _vp, _i64, _u64, _u32], C.c_int),
    "g2048_host_env_step": ([_vp] * 11 + [_i64, _u64, _u32], C.c_int),
    "g2048_host_env_rollout": ([_vp] * 6 + [_i64, _i32, _u32, _u64, _u32], C.c_int),
    "g2048_host_legal_masks": ([_vp, _vp, _vp, _i64], C.c_int),
    "g2048_host_evaluate": ([_vp, _vp, _vp, _i64], C.c_int),
    "g2048_host_beam_search": ([_vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _u64, _u32], C.c_int),
    "g2048_host_play_games": ([_i64, _i32, _i32, _i32, _i32, _i32, _u64, _u32] + [_vp] * 9, C.c_int),
}

EXPORTS = tuple(_SIGNATURES)


def load():
    """dlopen the library and declare every signature.  Does not touch the GPU."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise G2048Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `make -C {CSRC}` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        raise G2048Error(f"libg2048 error {rc}: {load().g2048_last_error().decode()}")


def use_device(index: int):
    """Initialises `index` on first use and makes it the calling thread's current CUDA device.

    The C ABI picks its per-device state from cudaGetDevice(), which is per THREAD and which torch
    changes behind our back (`torch.cuda.set_device`, `with torch.cuda.device(...)`, stream contexts),
    so the device is set on every entry instead of trusting a cache; cudaSetDevice is cheap."""
    lib = load()
    if index not in _ready_devices:
        check(lib.g2048_init(index))
        _ready_devices.add(index)
    else:
        check(lib.g2048_set_device(index))
    return lib


def launch_count() -> int:
    return int(load().g2048_launch_count())


def overflow_count(stream=None) -> int:
    v = _u64(0)
    check(load().g2048_overflow_count(C.byref(v), stream))
    return int(v.value)


def np_ptr(a):
    """void* of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return a.ctypes.data_as(C.c_void_p)
