"""ctypes binding of the C ABI in ``include/recboard_b200.h``.

PyTorch is used only for device memory and streams: every call passes raw device pointers
(``tensor.data_ptr()``) and the current CUDA stream.  There is NO CPU fallback -- if the shared
library is missing, or the device is not an sm_100 part, calls raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_SO = Path(__file__).resolve().parent / "_C" / ("librecboard_b200" + os.environ.get("RB_SO_SUFFIX", "") + ".so")

DTYPE_F32, DTYPE_BF16 = 0, 1
MODE_BF16, MODE_FP32X3 = 0, 1
OP_SCATTER_ADD, OP_SCORE_DENSE, OP_CE_FWD, OP_CE_BWD, OP_TOPK_EVAL = range(5)
MASKED_SCORE = -1e23

_p, _i64, _i32, _f32, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t

#: every symbol include/recboard_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "rb_gather_rows": (_i32, [_p, _p, _p, _i64, _i64, _i32, _i32, _p]),
    "rb_ipc_export": (_i32, [_p, _p, _p]),
    "rb_ipc_open": (_i32, [_p, _i64, _p, _p]),
    "rb_ipc_close": (_i32, [_p]),
    "rb_gather_rows_peers": (_i32, [_p, _p, _i32, _p, _p, _i64, _i32, _i32, _p]),
    "rb_compact_index": (_i32, [_p, _i64, _p, _p, _p]),
    "rb_normalize_rows": (_i32, [_p, _p, _p, _i64, _i32, _i32, _i32, _f32, _p]),
    "rb_scatter_add_rows": (_i32, [_p, _p, _p, _i64, _i64, _i32, _i32, _i64, _p, _sz, _p]),
    "rb_scatter_add_rows_into": (_i32, [_p, _p, _p, _i64, _i64, _i32, _i32, _i32, _i64, _p, _sz, _p]),
    "rb_gather_dot": (_i32, [_p, _p, _p, _f32, _p, _i64, _i64, _i64, _i32, _i32, _p]),
    "rb_gather_dot_bwd": (_i32, [_p, _p, _p, _p, _f32, _p, _p, _i64, _i64, _i64, _i32, _i32, _i64, _p, _sz, _p]),
    "rb_spmm_csr": (_i32, [_p, _p, _p, _p, _p, _p, _f32, _i64, _i64, _i32, _p]),
    "rb_score_dense": (_i32, [_p, _p, _p, _f32, _p, _i64, _i64, _i32, _i32, _i32, _p, _sz, _p]),
    "rb_ce_fwd": (_i32, [_p, _p, _p, _f32, _p, _i64, _i64, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "rb_ce_du_finish": (_i32, [_p, _p, _p, _p, _p, _i64, _f32, _f32, _p, _i64, _i64, _i32, _i32, _p, _p, _p]),
    "rb_ce_bwd": (_i32, [_p, _p, _p, _f32, _p, _i64, _p, _f32, _p, _i64, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _p, _sz, _p]),
    "rb_ce_bwd_dw_bf16": (_i32, [_p, _p, _p, _f32, _p, _i64, _p, _f32, _p, _i64, _i64, _i32, _p, _p, _p, _p, _sz, _p]),
    "rb_ce_bwd_dw_bf16_acc": (_i32, [_p, _p, _p, _f32, _p, _i64, _p, _f32, _p, _i64, _i64, _i32, _p, _p, _p, _p, _sz, _p]),
    "rb_topk_eval": (_i32, [_p, _p, _p, _f32, _p, _p, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _p, _p, _p, _sz, _p]),
    "rb_topk_debug_layout": (_i32, [_i64, _i64, _i32, _i32, _i32, _i64, _p]),
    "rb_topk_hits": (_i32, [_p, _p, _p, _i64, _i32, _p, _p]),
    "rb_rowstats_merge": (_i32, [_p, _i32, _i64, _p, _p, _p]),
    "rb_topk_merge_packed": (_i32, [_p, _i32, _i64, _i32, _p, _p, _p]),
    "rb_topk_metrics_blocks": (_i32, [_i64]),
    "rb_topk_metrics": (_i32, [_p, _p, _p, _i64, _i32, _p, _p, _p, _p, _i32, _p, _p, _p]),
    "rb_topk_merge": (_i32, [_p, _p, _i32, _i64, _i32, _p, _p, _p]),
    "rb_workspace_bytes": (_sz, [_i32, _i64, _i64, _i32, _i32, _i32, _i64]),
    "rb_launch_count": (_i64, []),
    "rb_last_error": (C.c_char_p, []),
    "rb_version": (C.c_char_p, []),
}

_lib = None


def lib() -> C.CDLL:
    """Load (once) the in-tree CUDA library; fail loudly when it is absent."""
    global _lib
    if _lib is None:
        if not _SO.exists():
            raise RuntimeError(
                f"{_SO} is missing: build it with `python -m recboard_b200.build` "
                "(recboard_b200 has no CPU / PyTorch fallback)"
            )
        h = C.CDLL(str(_SO))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


E_UNSUPPORTED = -3


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().rb_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {code}): {msg}")


def ptr(t):
    """Raw device pointer of a tensor handed to the C ABI (dense row-major layout is part of the contract)."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise RuntimeError("recboard_b200: a non-contiguous tensor reached the C ABI (call .contiguous() first)")
    return C.c_void_p(t.data_ptr())


def call(device, name: str, *args) -> None:
    """One C-ABI call with ``device`` made current for its duration: the library reads the SM count, sets
    function attributes and launches on the *current* device, while the pointers and the stream belong to
    the tensors' device -- the two must agree also when a process drives several GPUs."""
    with torch.cuda.device(device):
        check(getattr(lib(), name)(*args), name)


def stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return DTYPE_F32
    if t.dtype == torch.bfloat16:
        return DTYPE_BF16
    raise TypeError(f"unsupported dtype {t.dtype}: the path computes in bf16 or fp32")


def require_cuda(*tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "recboard_b200 ops run on a CUDA device only (no CPU fallback); got a CPU tensor"
            )
        dev = t.device if dev is None else dev
        if t.device != dev:
            raise RuntimeError("all tensors must live on the same device")
    return dev


class Workspace:
    """Grow-only scratch buffer handed to the C ABI (the library never allocates), one per (device, stream):
    calls on the same stream are ordered and may share it, calls on different streams must not."""

    _bufs = {}
    MAX_STREAMS = 8   # buffers kept per device; the least recently used one goes first (short-lived side streams)

    @classmethod
    def get(cls, device: torch.device, nbytes: int) -> torch.Tensor:
        if device.index is None:
            raise RuntimeError("recboard_b200: tensors must carry an explicit CUDA device index")
        key = (device.index, torch.cuda.current_stream(device).cuda_stream)
        buf = cls._bufs.pop(key, None)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        cls._bufs[key] = buf   # re-inserted last: dict order is the LRU order
        same_dev = [k for k in cls._bufs if k[0] == device.index]
        for k in same_dev[:-cls.MAX_STREAMS]:
            del cls._bufs[k]   # the caching allocator keeps the block alive until queued work on it has run
        return buf


def workspace_bytes(op: int, M: int, N: int, d: int, K: int = 0, mode: int = MODE_BF16, nnz: int = 0,
                    device=None) -> int:
    if device is None:
        return int(lib().rb_workspace_bytes(op, M, N, d, K, mode, nnz))
    with torch.cuda.device(device):   # the plan depends on the SM count of the device the call will run on
        return int(lib().rb_workspace_bytes(op, M, N, d, K, mode, nnz))


def launch_count() -> int:
    return int(lib().rb_launch_count())
