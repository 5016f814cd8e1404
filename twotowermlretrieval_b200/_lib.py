"""ctypes binding of libttr_b200.so (include/ttr_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C csrc`.  There is NO
CPU fallback: if the shared object is missing, or a call fails, we raise.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libttr_b200.so"

P, I32, I64, F32, F64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double

# name -> argument ctypes (the trailing stream pointer included)
_SIGNATURES = {
    "ttr_sm_count": [P],
    "ttr_debug_set_flags": [I32],
    "ttr_debug_get_flags": [P],
    "ttr_debug_gru_tc_max_clusters": [P],
    "ttr_debug_set_trace": [P],
    "ttr_pack_padded_i64": [P, P, P, P, I64, I64, P],
    "ttr_pack_padded_count_i64": [P, P, P, P, I64, I64, P, P, P],
    "ttr_seq_plan": [P, I32, I32, P, P, P, P, P],
    "ttr_embed_gather": [P, I32, I32, P, I64, I32, P, P, P, I32, P],
    "ttr_embed_scatter_grad": [P, I32, I32, I64, I32, P, P, P, P, P],
    "ttr_gemm_tf32_bias": [P, P, P, P, I32, P, I32, I32, P],
    "ttr_debug_gemm_fp32_bias": [P, P, P, P, I32, P, I32, I32, P],
    "ttr_gemm_tn_fp32": [P, P, P, I32, P, I32, I32, I32, P],
    "ttr_gemm_nn_fp32": [P, P, P, I32, P, I32, I32, I32, P],
    "ttr_gemm_tn_tf32": [P, I32, P, I32, P, I32, I32, P, I32, I32, I32, P],
    "ttr_zero_tail_rows": [P, I32, P, I32, P],
    "ttr_gru_recurrence_fwd": [P, P, P, P, P, I32, I32, I32, P, P, P, P],
    "ttr_gru_recurrence_fwd_ws": [P, P, P, P, P, I32, I32, I32, P, P, P, P, I64, P],
    "ttr_gru_recurrence_fwd_f16": [P, P, P, P, P, I32, I32, I32, P, P, P, I64, P],
    "ttr_gemm_f16_bias": [P, P, P, P, I32, P, I32, I32, P],
    "ttr_f32_to_f16": [P, P, I64, P],
    "ttr_gru_recurrence_bwd_ws": [P, P, P, P, P, P, P, I32, I32, I32, P, P, P, I64, I32, P, P, P, P],
    "ttr_gru_recurrence_bwd": [P, P, P, P, P, P, P, I32, I32, I32, P, P, P],
    "ttr_gru_whh_grad": [P, P, P, I32, I32, I32, I32, P, P, I32, P],
    "ttr_proj_l2norm_fwd": [P, P, P, I32, I32, I32, I32, P, P, P],
    "ttr_l2norm_bwd": [P, P, I32, I32, I32, P, P],
    "ttr_triplet_fwd": [P, P, P, I32, I32, F32, P, P],
    "ttr_triplet_bwd": [P, P, P, I32, I32, F32, P, P, P, P, P],
    "ttr_batch_metrics": [P, P, P, I32, I32, P, P],
    "ttr_colsum": [P, I32, P, I32, P, I32, P],
    "ttr_clip_adam": [P, P, P, P, I64, F32, F32, F32, F32, F32, F32, I32, P, P, P],
    "ttr_score_topk": [P, I32, P, I64, I32, I32, I64, P, P, P, I64, P],
    "ttr_topk_merge": [P, P, I32, I32, I32, I32, P, P, P],
    "ttr_topk_merge_peers": [P, P, P, I32, I32, I32, I32, P, P, P, P],
    "ttr_topk_exchange_merge": [P, P, P, P, I32, I32, ctypes.c_uint32, I32, I32, I32, P, P, P, P],
    "ttr_positive_rank": [P, P, P, I32, I64, I32, P, P, P],
    "ttr_dropout": [P, I64, F32, ctypes.c_uint64, P, P, P],
    "ttr_blend_topk": [P, F32, P, I64, I32, P, P, P, P, P, I32, F64, I32, P, P, P, P, P],
    "ttr_hybrid_rerank": [P, P, I32, I32, I64, P, P, P, P, P, P, P, P, P, F64, I32, I32, P, P, P, P, P],
    "ttr_tfidf_candidates": [P, I32, I32, I64, I64, P, P, P, P, P, P, P, P],
}

EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + ["ttr_last_error", "ttr_version", "ttr_score_topk_workspace_bytes",
                                               "ttr_blend_topk_workspace_bytes", "ttr_gru_fwd_workspace_bytes",
                                               "ttr_gru_bwd_workspace_bytes"])

_lib = None


class TTRError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the shared library (no CUDA call is made here)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("TTR_B200_LIB", LIB_PATH))
    if not path.exists():
        raise TTRError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = ctypes.CDLL(str(path))
    lib.ttr_last_error.restype = ctypes.c_char_p
    lib.ttr_last_error.argtypes = []
    lib.ttr_version.restype = ctypes.c_int
    lib.ttr_score_topk_workspace_bytes.restype = ctypes.c_int64
    lib.ttr_score_topk_workspace_bytes.argtypes = [I32, I64, I32]
    lib.ttr_blend_topk_workspace_bytes.restype = ctypes.c_int64
    lib.ttr_blend_topk_workspace_bytes.argtypes = [I32]
    lib.ttr_gru_fwd_workspace_bytes.restype = ctypes.c_int64
    lib.ttr_gru_fwd_workspace_bytes.argtypes = [I32, I32, I32]
    lib.ttr_gru_bwd_workspace_bytes.restype = ctypes.c_int64
    lib.ttr_gru_bwd_workspace_bytes.argtypes = [I32, I32, I32]
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(t):
    """Device (or host) address of a tensor, None -> NULL."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _owning_device(args):
    """Device of the CUDA tensor arguments of one call (None when there is none); raises when
    they live on different devices — the library launches on ONE device and stream."""
    dev = None
    for a in args:
        if isinstance(a, torch.Tensor) and a.is_cuda:
            if dev is None:
                dev = a.device
            elif a.device != dev:
                raise TTRError(f"CUDA tensor arguments on different devices ({dev} and {a.device})")
    return dev


def call(name: str, *args):
    """Invoke `name`; tensors are passed by address.  The launch happens on the device that owns the
    tensor arguments (not the thread's current device) and on torch's current stream OF THAT DEVICE,
    which is appended as the last argument."""
    lib = load()
    dev = _owning_device(args)
    conv = [ptr(a) if isinstance(a, torch.Tensor) else a for a in args]
    fn = getattr(lib, name)
    if dev is None or dev.index == torch.cuda.current_device():
        rc = fn(*conv, stream_ptr())
    else:
        with torch.cuda.device(dev):
            rc = fn(*conv, stream_ptr(dev))
    if rc != 0:
        raise TTRError(f"{name} failed ({rc}): {lib.ttr_last_error().decode()}")


def call_nostream(name: str, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise TTRError(f"{name} failed ({rc}): {lib.ttr_last_error().decode()}")


_SM_COUNT = {}


def sm_count(device=None) -> int:
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    if dev is None:
        dev = torch.cuda.current_device()
    if dev not in _SM_COUNT:
        out = ctypes.c_int(0)
        with torch.cuda.device(dev):
            call_nostream("ttr_sm_count", ctypes.byref(out))
        _SM_COUNT[dev] = out.value
    return _SM_COUNT[dev]


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise TTRError(f"{what}: expected a CUDA tensor; twotowermlretrieval_b200 has no CPU path")
