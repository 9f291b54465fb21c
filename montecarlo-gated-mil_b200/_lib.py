"""ctypes binding of include/mcmil_b200.h (the C-ABI of the CUDA library).

The product path has no fallback: if `lib/libmcmil_b200.so` is missing or does not export a
symbol, loading raises.  (`build.py` / `__graft_entry__.build()` produce the library.)
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
# MCMIL_LIB_PATH: A/B experiments with alternative builds of the same C ABI (tools/build_variants.py)
LIB_PATH = os.environ.get("MCMIL_LIB_PATH") or os.path.join(HERE, "lib", "libmcmil_b200.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "mcmil_b200.h")

IMPL_TCGEN05 = 0
IMPL_SIMT_FP32 = 1
IMPLS = {"tcgen05": IMPL_TCGEN05, "simt_fp32": IMPL_SIMT_FP32}

_vp, _i, _f, _u64, _sz, _dbl = C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_size_t, C.c_double

# name -> (restype, argtypes); must cover every function include/mcmil_b200.h declares
SIGNATURES = {
    "mcmil_last_error": (C.c_char_p, []),
    "mcmil_version": (_i, []),
    "mcmil_weights_create": (_i, [C.POINTER(_vp), _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mcmil_weights_destroy": (_i, [_vp]),
    "mcmil_plan_create": (_i, [C.POINTER(_vp), C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i, _i, _i, _vp]),
    "mcmil_plan_destroy": (_i, [_vp]),
    "mcmil_plan_workspace_bytes": (_sz, [_vp]),
    "mcmil_plan_total_rows": (_i, [_vp]),
    "mcmil_plan_plane_cols": (_i, [_vp]),
    "mcmil_plan_set_sm_limit": (_i, [_vp, _i]),
    "mcmil_plan_bag_plane_col": (_i, [_vp, _i]),
    "mcmil_head_forward": (_i, [_vp, _vp, _vp, _i, _i, _u64, _i, _f, _f, _vp, _vp, _i,
                                _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mcmil_head_forward_f16": (_i, [_vp, _vp, _vp, _i, _i, _u64, _i, _f, _f, _vp, _vp, _i,
                                    _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mcmil_welford_pack": (_i, [_vp, _vp, _dbl, _i, _vp, _vp]),
    "mcmil_welford_unpack": (_i, [_vp, _i, _vp, _vp, _vp]),
    "mcmil_export_masks": (_i, [_vp, _i, _i, _u64, _i, _f, _f, _vp, _vp, _vp]),
    "mcmil_debug_proj_tc": (_i, [_vp, _vp, _vp, _i, _i, _u64, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mcmil_last_launch_count": (_i, []),
    "mcmil_attnmap_stats": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "mcmil_tile_nonzero_pct": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp]),
    "mcmil_gather_tiles": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "mcmil_aux_pairwise_loss": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _f, _f, _vp, _vp]),
    "mcmil_profile_begin": (_i, [_i]),
    "mcmil_profile_end": (_i, [C.POINTER(_dbl), C.POINTER(_i)]),
}

_lib = None


def declared_symbols() -> list[str]:
    """Function names declared in include/mcmil_b200.h."""
    with open(HEADER_PATH) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(mcmil_[a-z0-9_]+)\s*\(", text)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the B200 CUDA library is not built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(code: int, what: str):
    if code == 0:
        return
    msg = load().mcmil_last_error().decode()
    if code in (-1, -3):
        raise ValueError(f"{what}: {msg}")
    if code == -4:
        raise NotImplementedError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg} (code {code})")
