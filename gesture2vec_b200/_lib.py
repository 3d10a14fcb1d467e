"""ctypes binding of the C-ABI CUDA library (include/g2v_vq.h).

There is no CPU fallback: if the library is missing, or an entry point fails, a
RuntimeError is raised.  Build it with ``python -m gesture2vec_b200.build`` (or
``__graft_entry__.build()``); the .so lives in-tree at csrc/libg2v_vq.so.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# G2V_LIB_PATH selects an experiment build of the same library (python -m gesture2vec_b200.build --suffix=...)
LIB_PATH = os.environ.get("G2V_LIB_PATH") or os.path.join(HERE, "csrc", "libg2v_vq.so")

# dtype / flag codes (mirror include/g2v_vq.h)
F32, BF16, F16 = 0, 1, 2
ALGO_AUTO, ALGO_SIMT, ALGO_TC, NO_RECHECK, NO_REFINE, LIST_ALL_ROWS = 0, 1, 2, 4, 8, 16
ALGO_MASK = 3
ERR_UNSUPPORTED = -7
GEMM_ACCUMULATE, GEMM_FP16 = 1, 2
DET_CHUNK = 128                    # G2V_DET_CHUNK
TC_VARIANT_TMEM, TC_VARIANT_FUSED, TC_VARIANT_PREP = 1 << 8, 2 << 8, 3 << 8
STAT_ROWS, STAT_PAIR_RECHECK, STAT_FULL_RECHECK, STAT_FALLBACK_ROWS, STAT_REFINE_ROWS, STAT_REFINE_EXACT = 0, 1, 2, 3, 4, 5

_p, _i, _i64, _f, _sz, _u = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t, C.c_uint

# name -> (restype, argtypes); every symbol include/g2v_vq.h declares
SIGNATURES = {
    "g2v_version": (_i, []),
    "g2v_strerror": (C.c_char_p, [_i]),
    "g2v_last_error_detail": (C.c_char_p, []),
    "g2v_launch_count": (C.c_ulonglong, []),
    "g2v_codebook_bytes": (_sz, [_i, _i]),
    "g2v_codebook_prepare": (_i, [_p, _i, _i, _p, _sz, _p]),
    "g2v_profile_next_search": (_i, [_p, _p]),
    "g2v_search_path": (_i, [_i, _i, _u]),
    "g2v_workspace_bytes": (_sz, [_i64, _i, _i, _i, _u]),
    "g2v_vq_search": (_i, [_p, _i, _p, _p, _i64, _i, _i, _p, _p, _p, _sz, _u, _p]),
    "g2v_vq_search_wide": (_i, [_p, _i, _i, _p, _p, _i64, _i, _i, _p, _p, _p, _sz, _u, _p]),
    "g2v_vq_apply": (_i, [_p, _p, _p, _p, _i64, _i, _i, _p, _p, _p, _p, _i, _p]),
    "g2v_apply_workspace_bytes": (_sz, [_i64, _i]),
    "g2v_vq_apply_ws": (_i, [_p, _p, _p, _p, _i64, _i, _i, _p, _p, _p, _p, _i, _p, _sz, _p]),
    "g2v_fold_projection": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "g2v_pad_rows": (_i, [_p, _i64, _i, _i, _p, _p]),
    "g2v_vq_stats_deterministic": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i, _i, _p, _p, _p, _p, _p]),
    "g2v_vq_stats_pack": (_i, [_p, _p, _p, _i, _i64, _i, _i, _p, _p]),
    "g2v_vq_stats_finalize": (_i, [_p, _i, _i, _f, _f, _p, _p, _p]),
    "g2v_vq_ema_update": (_i, [_p, _p, _p, _p, _p, _p, _p, _f, _f, _i, _i, _p, _sz, _p]),
    "g2v_vq_step_finalize": (_i, [_p, _p, _p, _i, _i64, _p, _i, _i, _f, _f, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _f, _f,
                                  _p, _p, _sz, _p]),
    "g2v_gemm_workspace_bytes": (_sz, [_i64, _i, _i64, _u]),
    "g2v_gemm_f32": (_i, [_p, _i64, _i, _p, _i64, _i, _i64, _i, _i64, _p, _p, _i64, _f, _u, _p, _sz, _p]),
    "g2v_soft_assign": (_i, [_p, _p, _p, _p, _i64, _i, _i, _p, _p, _p]),
    "g2v_soft_tail": (_i, [_p, _p, _i64, _i, _i, _f, _p, _p, _p, _p, _p, _p]),
    "g2v_soft_backward": (_i, [_p, _p, _p, _p, _i64, _i, _p, _p, _p, _p, _p, _p]),
    "g2v_soft_gx": (_i, [_p, _p, _p, _p, _p, _p, _i64, _p, _p]),
    "g2v_exact_workspace_bytes": (_sz, [_i]),
    "g2v_vq_search_exact": (_i, [_p, _i, _p, _i64, _i, _i, _p, _p, _sz, _p]),
    "g2v_vq_backward": (_i, [_p, _p, _p, _p, _p, _f, _i64, _i, _i, _p, _p]),
    "g2v_vq_grad_codebook": (_i, [_p, _p, _f, _i, _i, _p, _p]),
    "g2v_kmeans_update": (_i, [_p, _p, _i, _i, _p, _p, _p, _sz, _p]),
    "g2v_onehot": (_i, [_p, _i64, _i, _p, _p]),
    "g2v_tokenize_host_bytes": (_sz, [_i64, _i, _i, _i, _u]),
    "g2v_tokenize_host": (_i, [_p, _i, _i64, _p, _p, _i, _i, _p, _i64, _p, _p, _sz, _u]),
}

_lib = None


def load() -> C.CDLL:
    """Load the library once; raise loudly if it is absent or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"gesture2vec_b200: CUDA library not found at {LIB_PATH}. Build it with "
            "`python -m gesture2vec_b200.build`. There is no CPU fallback for the quantizer path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover
            raise RuntimeError(f"gesture2vec_b200: {LIB_PATH} does not export {name}") from e
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        lib = load()
        msg = lib.g2v_strerror(rc).decode()
        detail = lib.g2v_last_error_detail().decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc}){': ' + detail if detail else ''}")
