"""Host-side operators of the full-catalog scoring path (thin wrappers over the C ABI).

Each function names the reference lines it stands in for (paths relative to the reference
root).  All tensors must be CUDA tensors: there is no CPU fallback.

Precision: ``precision="bf16"`` runs the contraction on bf16 operands with fp32 accumulation
(configs 3-5); ``precision="fp32"`` runs the error-compensated 3xTF32 tensor-core mode whose
scores agree with an fp32 SGEMM to ~1e-6 (configs 1-2).  Default: taken from the operand dtype.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _lib as L


#: RB_CHECK_FINITE=1 (or ``ops.CHECK_FINITE = True``): every fused_ce forward verifies that the row log-sum-exps are
#: finite and raises FloatingPointError otherwise.  Off by default: the check reads one flag back from the device.
CHECK_FINITE = os.environ.get("RB_CHECK_FINITE", "0") not in ("", "0")


def _mode_for(U: torch.Tensor, precision: Optional[str]) -> str:
    if precision is None:
        precision = "bf16" if U.dtype == torch.bfloat16 else "fp32"
    if precision not in ("bf16", "fp32"):
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
    return precision


def _pad_cols(x: torch.Tensor, mult: int = 8) -> torch.Tensor:
    """Zero columns up to a multiple of ``mult``: the kernels move rows as 16-byte vectors (TMA rows, 128-bit loads), and
    zero columns change no dot product or norm.  For embedding widths like HSTU's 50
    (HSTU/configs/MovieLens1M_500_LOU.yaml) this costs a padded copy of the operand per call; differentiable
    (autograd slices the gradient back)."""
    r = x.shape[-1] % mult
    return x if r == 0 else torch.nn.functional.pad(x, (0, mult - r))


def _prep(U, W, precision):
    """Cast operands to the storage type the chosen arithmetic mode needs (rows padded to whole 16-byte vectors)."""
    precision = _mode_for(U, precision)
    U, W = _pad_cols(U), _pad_cols(W)
    if precision == "bf16":
        return U.to(torch.bfloat16).contiguous(), W.to(torch.bfloat16).contiguous(), L.MODE_BF16
    return U.float().contiguous(), W.float().contiguous(), L.MODE_FP32X3


def _ws(dev, op, M, N, d, K=0, mode=L.MODE_BF16, nnz=0):
    n = L.workspace_bytes(op, M, N, d, K, mode, nnz, device=dev)
    return L.Workspace.get(dev, n), n


# --------------------------------------------------------------------------------------
# item-embedding gather + deterministic scatter-add backward
# --------------------------------------------------------------------------------------
def gather_rows_raw(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    dev = L.require_cuda(table, idx)
    if idx.dtype != torch.int64:
        raise TypeError("idx must be int64")
    table, idx = table.contiguous(), idx.contiguous()
    n_rows, d = table.shape
    out = torch.empty(*idx.shape, d, dtype=table.dtype, device=dev)
    L.call(dev, "rb_gather_rows", L.ptr(table), L.ptr(idx), L.ptr(out), idx.numel(), n_rows, d,
                               L.dtype_code(table), L.stream_ptr(dev))
    return out


def scatter_add_rows_(grad_table: torch.Tensor, grad_out: torch.Tensor, idx: torch.Tensor,
                      padding_idx: int = -1) -> torch.Tensor:
    """grad_table[idx[i]] += grad_out[i] in place (fp32 accumulate, fixed summation order); ``grad_table`` is fp32
    or bf16 (one rounding per touched row).  One cooperative launch (csrc/scatter.cuh)."""
    dev = L.require_cuda(grad_table, grad_out, idx)
    if grad_table.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("grad_table must be float32 or bfloat16")
    n_rows, d = grad_table.shape
    n_idx = idx.numel()
    grad_out, idx = grad_out.contiguous(), idx.contiguous()   # grad_table is updated in place: it must be dense already
    ws, n = _ws(dev, L.OP_SCATTER_ADD, 0, 0, d, nnz=n_idx)
    L.call(dev, "rb_scatter_add_rows_into", L.ptr(grad_out), L.ptr(idx), L.ptr(grad_table), n_idx, n_rows, d,
           L.dtype_code(grad_out), L.dtype_code(grad_table), padding_idx, L.ptr(ws), n, L.stream_ptr(dev))
    return grad_table


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, idx, padding_idx, accumulate):
        ctx.save_for_backward(idx)
        ctx.shape, ctx.dtype, ctx.padding_idx = table.shape, table.dtype, padding_idx
        # accumulate=True: the backward adds straight into ``table.grad`` (a leaf's persistent gradient buffer)
        ctx.leaf = table if (accumulate and table.is_leaf and table.requires_grad) else None
        return gather_rows_raw(table.contiguous(), idx.contiguous())

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        go = grad_out.contiguous()
        if go.dtype not in (torch.float32, torch.bfloat16):
            go = go.float()
        go, flat = go.view(-1, ctx.shape[1]), idx.contiguous().view(-1)
        leaf = ctx.leaf
        if leaf is not None and leaf.grad is not None and leaf.grad.is_contiguous() and leaf.grad.shape == ctx.shape \
                and leaf.grad.dtype in (torch.float32, torch.bfloat16):
            # the table's existing gradient gets the rows added in place (it may already hold, or later receive, the
            # dW of the scoring head: the reference's autograd sums both into this one (N+P,d) tensor too)
            scatter_add_rows_(leaf.grad, go, flat, ctx.padding_idx)
            return None, None, None, None
        gdt = ctx.dtype if ctx.dtype in (torch.float32, torch.bfloat16) else torch.float32
        g = torch.zeros(ctx.shape, dtype=gdt, device=grad_out.device)   # in the parameter's dtype: no fp32 copy + cast
        scatter_add_rows_(g, go, flat, ctx.padding_idx)
        return g.to(ctx.dtype), None, None, None


def gather_rows(table: torch.Tensor, idx: torch.Tensor, padding_idx: int = -1, accumulate: bool = False) -> torch.Tensor:
    """``self.Item.embeddings(seqs)`` (SASRec/main.py:183) with the ``nn.Embedding(padding_idx=)``
    backward contract: dense table gradient, ``padding_idx`` row left at zero.

    ``accumulate=True`` (training loops that keep ``table.grad`` allocated, ``zero_grad(set_to_none=False)``): the
    backward adds the rows directly into ``table.grad`` instead of returning a fresh dense (N+P,d) tensor for
    autograd to add -- the zero-fill and the extra read-modify-write of the whole table disappear.  Not for
    ``torch.autograd.grad`` / double backward (the side effect is on ``.grad``)."""
    return _GatherRows.apply(table, idx, padding_idx, accumulate)


# --------------------------------------------------------------------------------------
# device-side compaction of query rows (a4): no nonzero(), no count read back to the host
# --------------------------------------------------------------------------------------
def compact_index(mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(row_index int64 (n,), count int32 (1,)) of a boolean mask with n elements: row_index[k] is the flat position of
    the k-th True for k < count and -1 beyond; both stay on the device (rb_compact_index)."""
    dev = L.require_cuda(mask)
    m8 = mask.reshape(-1).to(torch.uint8).contiguous()
    n = m8.numel()
    row_index = torch.empty(n, dtype=torch.int64, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    L.call(dev, "rb_compact_index", L.ptr(m8), n, L.ptr(row_index), L.ptr(count), L.stream_ptr(dev))
    return row_index, count


def compact_queries(X: torch.Tensor, mask: torch.Tensor, *aligned: torch.Tensor):
    """The sync-free stand-in for ``X[mask]`` / ``t[mask]`` (``userEmbds[indices]``, ``positives[indices]``,
    SASRec/main.py:199-200): -> (Xc (n, d), [t_c (n,) ...], count) where n = mask.numel() is the CAPACITY, the first
    ``count`` rows are the selected ones in order, the rest are zero rows / -1 entries.  Hand ``count`` to ``fused_ce`` as
    ``n_valid``.  Differentiable w.r.t. ``X`` (the backward scatters the rows back; the selected positions are unique)."""
    d = X.shape[-1]
    row_index, count = compact_index(mask)
    Xc = gather_rows(X.reshape(-1, d), row_index, padding_idx=-1)
    safe = row_index.clamp_min(0)
    outs = [torch.where(row_index >= 0, t.reshape(-1)[safe], torch.full_like(safe, -1).to(t.dtype)) for t in aligned]
    return Xc, outs, count


# --------------------------------------------------------------------------------------
# fused gather + row-wise dot (pool ranking, sampled softmax, BPR / BCE logits)
# --------------------------------------------------------------------------------------
def _same_storage(U, table):
    dt = torch.bfloat16 if (U.dtype == torch.bfloat16 and table.dtype == torch.bfloat16) else torch.float32
    return U.detach().to(dt).contiguous(), table.detach().to(dt).contiguous()


class _GatherDot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, U, table, idx, scale, padding_idx):
        dev = L.require_cuda(U, table, idx)
        if idx.dtype != torch.int64 or idx.dim() != 2 or idx.shape[0] != U.shape[0]:
            raise TypeError("idx must be int64 of shape (M, K)")
        Uc, Tc = _same_storage(U, table)
        idc = idx.contiguous()
        M, d = Uc.shape
        K = idc.shape[1]
        S = torch.empty(M, K, dtype=torch.float32, device=dev)
        L.call(dev, "rb_gather_dot", L.ptr(Uc), L.ptr(Tc), L.ptr(idc), float(scale), L.ptr(S), M, K, Tc.shape[0], d,
                                  L.dtype_code(Uc), L.stream_ptr(dev))
        ctx.save_for_backward(Uc, Tc, idc)
        ctx.scale, ctx.padding_idx, ctx.dtypes = float(scale), int(padding_idx), (U.dtype, table.dtype)
        return S

    @staticmethod
    def backward(ctx, G):
        Uc, Tc, idc = ctx.saved_tensors
        dev = Uc.device
        M, d = Uc.shape
        K = idc.shape[1]
        need_u, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dU = torch.empty(M, d, dtype=torch.float32, device=dev) if need_u else None
        dT = torch.zeros(Tc.shape[0], d, dtype=torch.float32, device=dev) if need_t else None
        ws, n = _ws(dev, L.OP_SCATTER_ADD, 0, 0, d, nnz=M * K)
        L.call(dev, "rb_gather_dot_bwd", L.ptr(Uc), L.ptr(Tc), L.ptr(idc), L.ptr(G.float().contiguous()), ctx.scale, L.ptr(dU),
                                      L.ptr(dT), M, K, Tc.shape[0], d, L.dtype_code(Uc), ctx.padding_idx, L.ptr(ws), n,
                                      L.stream_ptr(dev))
        return (dU.to(ctx.dtypes[0]) if need_u else None, dT.to(ctx.dtypes[1]) if need_t else None, None, None, None)


def gather_dot(U: torch.Tensor, table: torch.Tensor, idx: torch.Tensor, scale: float = 1.0, padding_idx: int = -1) -> torch.Tensor:
    """``torch.einsum("MD,MKD->MK", U, table[idx]) * scale`` -> fp32 (M,K) without the (M,K,D) gather
    (recommend_from_pool SASRec/main.py:230-236; sampled softmax HSTU/main.py:192-197; BPR/BCE logits
    SASRec/main.py:203-206).  Differentiable w.r.t. ``U`` and ``table`` (dense, deterministic)."""
    if U.shape[-1] % 8:
        U, table = _pad_cols(U), _pad_cols(table)
    return _GatherDot.apply(U, table, idx, scale, padding_idx)


# --------------------------------------------------------------------------------------
# CSR x dense (LightGCN propagation)
# --------------------------------------------------------------------------------------
def _csr_parts(A: torch.Tensor):
    if A.layout != torch.sparse_csr:
        A = A.to_sparse_csr()
    return (A.crow_indices().to(torch.int64).contiguous(), A.col_indices().to(torch.int64).contiguous(),
            A.values().float().contiguous(), A.shape)


def spmm_raw(A: torch.Tensor, X: torch.Tensor, acc: Optional[torch.Tensor] = None, beta: float = 1.0,
             want_y: bool = True) -> Optional[torch.Tensor]:
    """Y = A @ X (fp32) through rb_spmm_csr; with ``acc`` also ``acc += beta * Y`` in the same pass."""
    crow, col, val, shape = _csr_parts(A)
    dev = L.require_cuda(crow, col, val, X, acc)
    Xc = X.detach().float().contiguous()
    d = Xc.shape[1]
    Y = torch.empty(shape[0], d, dtype=torch.float32, device=dev) if want_y else None
    if val.numel() == 0:   # empty matrix: nothing to launch
        return Y.zero_() if want_y else None
    L.call(dev, "rb_spmm_csr", L.ptr(crow), L.ptr(col), L.ptr(val), L.ptr(Xc), L.ptr(Y), L.ptr(acc), float(beta), shape[0], shape[1],
                            d, L.stream_ptr(dev))
    return Y


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, At, X):
        ctx.At = At
        return spmm_raw(A, X)

    @staticmethod
    def backward(ctx, dY):
        return None, None, spmm_raw(ctx.At, dY.contiguous())


def spmm(A: torch.Tensor, X: torch.Tensor, symmetric: bool = False, At: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``A @ X`` for a sparse CSR ``A`` (LightGCN/main.py:83), differentiable w.r.t. ``X``.  The backward is
    ``A^T @ dY``: pass ``symmetric=True`` for the symmetric-normalised adjacency (no transpose needed) or a
    cached ``At``; otherwise the transpose is built here."""
    if At is None:
        At = A if symmetric else A.t().to_sparse_csr()
    return _SpMM.apply(A, At, X)


# --------------------------------------------------------------------------------------
# row normalisation (cosine scoring operands)
# --------------------------------------------------------------------------------------
def normalize_rows(x: torch.Tensor, out_dtype: Optional[torch.dtype] = None, eps: float = 1e-12,
                   return_inv_norm: bool = False):
    """``F.normalize(x, dim=-1)`` for a (rows, d) matrix (HSTU/main.py:180-184) in one HBM pass,
    optionally cast to ``out_dtype`` (bf16 operand copy for the cosine sweep).  No autograd: this is
    the evaluation-side copy built once per sweep."""
    dev = L.require_cuda(x)
    if x.dim() != 2:
        raise ValueError("normalize_rows expects a (rows, d) matrix")
    xc = x.detach()
    if xc.dtype not in (torch.float32, torch.bfloat16):
        xc = xc.float()
    d_true = xc.shape[1]
    xc = _pad_cols(xc).contiguous()
    out_dtype = out_dtype or xc.dtype
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("out_dtype must be float32 or bfloat16")
    n, d = xc.shape
    out = torch.empty(n, d, dtype=out_dtype, device=dev)
    inv = torch.empty(n, dtype=torch.float32, device=dev) if return_inv_norm else None
    L.call(dev, "rb_normalize_rows", L.ptr(xc), L.ptr(out), L.ptr(inv), n, d, L.dtype_code(xc), L.dtype_code(out),
                                  float(eps), L.stream_ptr(dev))
    if d != d_true:
        out = out[:, :d_true].contiguous()
    return (out, inv) if return_inv_norm else out


# --------------------------------------------------------------------------------------
# dense scores (compatibility path of recommend_from_full)
# --------------------------------------------------------------------------------------
def score_dense(U: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor] = None, scale: float = 1.0,
                precision: Optional[str] = None) -> torch.Tensor:
    """``torch.einsum("BD,ND->BN", U, W)`` (SASRec/main.py:228) -> fp32 (B,N); no autograd."""
    dev = L.require_cuda(U, W, bias)
    Uc, Wc, mode = _prep(U.detach(), W.detach(), precision)
    M, d = Uc.shape
    N = Wc.shape[0]
    b = None if bias is None else bias.detach().float().contiguous()
    S = torch.empty(M, N, dtype=torch.float32, device=dev)
    ws, n = _ws(dev, L.OP_SCORE_DENSE, M, N, d, mode=mode)
    L.call(dev, "rb_score_dense", L.ptr(Uc), L.ptr(Wc), L.ptr(b), float(scale), L.ptr(S), M, N, d,
                               L.dtype_code(Uc), mode, L.ptr(ws), n, L.stream_ptr(dev))
    return S


# --------------------------------------------------------------------------------------
# fused full-catalog cross-entropy
# --------------------------------------------------------------------------------------
def fused_du_supported(U, precision: Optional[str], scale: float) -> bool:
    """Can rb_ce_fwd also produce the dU accumulator in the same sweep?  (bf16 mode, d <= 256, scale > 0)"""
    return _mode_for(U, precision) == "bf16" and U.shape[1] <= 256 and scale > 0


def _count_ptr(n_valid, dev):
    if n_valid is None:
        return None
    if n_valid.dtype != torch.int32 or n_valid.numel() != 1 or n_valid.device != dev:
        raise TypeError("n_valid must be an int32 scalar tensor on the operands' device")
    return L.ptr(n_valid.reshape(1))


def rowstats_merge(stats: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """stats (R,3,M) = per-rank (max, sumexp, label_logit) of a row-sharded table -> (global lse (M,), label_logit (M,))
    in one launch (rb_rowstats_merge)."""
    dev = L.require_cuda(stats)
    R, three, M = stats.shape
    st = stats.float().contiguous()
    out = torch.empty(2, M, dtype=torch.float32, device=dev)
    L.call(dev, "rb_rowstats_merge", L.ptr(st), R, M, L.ptr(out[0]), L.ptr(out[1]), L.stream_ptr(dev))
    return out[0], out[1]


def ce_rowstats(U, W, labels, bias=None, scale: float = 1.0, label_base: int = 0,
                precision: Optional[str] = None, want_dU: bool = False, n_valid: Optional[torch.Tensor] = None):
    """(row_max, row_sumexp, label_logit) of scale*U W^T + bias over this shard; (M,N) never exists.
    With ``want_dU`` a fourth tensor is returned: dU_unnorm (M,d) = sum_j exp(S_ij - row_max_i) W_j,
    accumulated by the same sweep (see ``ce_du_finish``)."""
    dev = L.require_cuda(U, W, labels, bias)
    Uc, Wc, mode = _prep(U.detach(), W.detach(), precision)
    M, d = Uc.shape
    N = Wc.shape[0]
    if labels.dtype != torch.int64 or labels.numel() != M:
        raise TypeError("labels must be int64 of shape (M,)")
    b = None if bias is None else bias.detach().float().contiguous()
    out = torch.empty(3, M, dtype=torch.float32, device=dev)
    du = torch.empty(M, d, dtype=torch.float32, device=dev) if want_dU else None
    ws, n = _ws(dev, L.OP_CE_FWD, M, N, d, mode=mode)
    L.call(dev, "rb_ce_fwd", L.ptr(Uc), L.ptr(Wc), L.ptr(b), float(scale), L.ptr(labels.contiguous()), label_base,
                          M, N, d, L.dtype_code(Uc), mode, L.ptr(out[0]), L.ptr(out[1]), L.ptr(out[2]), L.ptr(du),
                          _count_ptr(n_valid, dev), L.ptr(ws), n, L.stream_ptr(dev))
    if want_dU:
        return out[0], out[1], out[2], du
    return out[0], out[1], out[2]


def ce_du_finish(du_unnorm, row_max, lse, W, labels, grad_scale: float, scale: float = 1.0, label_base: int = 0,
                 grad_scale_dev: Optional[torch.Tensor] = None, n_valid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """This shard's piece of dU: g*scale*(du_unnorm*exp(row_max - lse) - [label in shard] W[label])."""
    dev = L.require_cuda(du_unnorm, row_max, lse, W, labels, grad_scale_dev)
    Wc = W.detach()
    if Wc.dtype not in (torch.float32, torch.bfloat16):
        Wc = Wc.float()
    Wc = Wc.contiguous()
    M, d = du_unnorm.shape
    dU = torch.empty(M, d, dtype=torch.float32, device=dev)
    L.call(dev, "rb_ce_du_finish", L.ptr(du_unnorm), L.ptr(row_max), L.ptr(lse.float().contiguous()), L.ptr(Wc),
                                L.ptr(labels.contiguous()), label_base, float(scale), float(grad_scale),
                                L.ptr(grad_scale_dev), M, Wc.shape[0], d, L.dtype_code(Wc), L.ptr(dU),
                                _count_ptr(n_valid, dev), L.stream_ptr(dev))
    return dU


def ce_backward(U, W, labels, lse, grad_scale: float, bias=None, scale: float = 1.0, label_base: int = 0,
                need_dU: bool = True, need_dW: bool = True, need_dbias: bool = False,
                precision: Optional[str] = None, grad_scale_dev: Optional[torch.Tensor] = None,
                dw_dtype: Optional[torch.dtype] = None, dw_out: Optional[torch.Tensor] = None,
                accumulate: bool = False, n_valid: Optional[torch.Tensor] = None):
    """Gradients of ``g * sum_i (lse_i - S_i,label_i)`` -> (dU, dW, dbias) fp32, with
    ``g = grad_scale * grad_scale_dev`` (the latter an optional fp32 device scalar).
    ``dw_dtype=torch.bfloat16`` (bf16 mode) makes the pass store dW in bf16 itself -- the correctly rounded
    fp32 result, without the fp32 (N,d) matrix and the cast pass a bf16 parameter would otherwise need.
    ``dw_out``: a contiguous (N,d) tensor (e.g. rows [P:] of a full (N+P,d) gradient) that receives dW;
    with ``accumulate`` (bf16 ``dw_out`` only) the rows are ADDED to its content inside the pass -- raises
    ``NotImplementedError`` when the direct bf16 pass does not apply (the caller adds a separate dW then)."""
    dev = L.require_cuda(U, W, labels, lse, bias, grad_scale_dev)
    if grad_scale_dev is not None and (grad_scale_dev.dtype != torch.float32 or grad_scale_dev.numel() != 1):
        raise TypeError("grad_scale_dev must be a float32 scalar tensor")
    if U.shape[1] % 8:   # rows are not whole 16-byte vectors: run on zero-padded copies, hand back the true columns
        if dw_out is not None:
            raise TypeError("dw_out needs an embedding width that is a multiple of 8")
        d0 = U.shape[1]
        dU, dW, db = ce_backward(_pad_cols(U.detach()), _pad_cols(W.detach()), labels, lse, grad_scale, bias, scale, label_base,
                                 need_dU, need_dW, need_dbias, precision, grad_scale_dev, dw_dtype, None, False, n_valid)
        return (dU[:, :d0].contiguous() if dU is not None else None, dW[:, :d0].contiguous() if dW is not None else None, db)
    Uc, Wc, mode = _prep(U.detach(), W.detach(), precision)
    M, d = Uc.shape
    N = Wc.shape[0]
    b = None if bias is None else bias.detach().float().contiguous()
    dU = torch.empty(M, d, dtype=torch.float32, device=dev) if need_dU else None
    db = torch.empty(N, dtype=torch.float32, device=dev) if need_dbias else None
    ws, n = _ws(dev, L.OP_CE_BWD, M, N, d, mode=mode)
    if need_dW and dw_dtype == torch.bfloat16 and mode == L.MODE_BF16:
        if need_dU:   # dU alone through the generic entry, dW (+ dbias) through the bf16-output pass
            L.call(dev, "rb_ce_bwd", L.ptr(Uc), L.ptr(Wc), L.ptr(b), float(scale), L.ptr(labels.contiguous()), label_base,
                                  L.ptr(lse.float().contiguous()), float(grad_scale), L.ptr(grad_scale_dev), M, N, d,
                                  L.dtype_code(Uc), mode, L.ptr(dU), None, None, _count_ptr(n_valid, dev), L.ptr(ws), n,
                                  L.stream_ptr(dev))
        dWb = dw_out if dw_out is not None else torch.empty(N, d, dtype=torch.bfloat16, device=dev)
        if dWb.dtype != torch.bfloat16 or dWb.shape != (N, d):
            raise TypeError("dw_out must be a bfloat16 (N,d) tensor here")
        args = (L.ptr(Uc), L.ptr(Wc), L.ptr(b), float(scale), L.ptr(labels.contiguous()), label_base,
                L.ptr(lse.float().contiguous()), float(grad_scale), L.ptr(grad_scale_dev), M, N, d,
                L.ptr(dWb), L.ptr(db), _count_ptr(n_valid, dev), L.ptr(ws), n, L.stream_ptr(dev))
        if accumulate:
            with torch.cuda.device(dev):
                code = L.lib().rb_ce_bwd_dw_bf16_acc(*args)
            if code == L.E_UNSUPPORTED:
                raise NotImplementedError("in-place accumulation of dW is not available for this shape")
            L.check(code, "rb_ce_bwd_dw_bf16_acc")
        else:
            L.call(dev, "rb_ce_bwd_dw_bf16", *args)
        return dU, dWb, db
    if accumulate:
        raise NotImplementedError("in-place accumulation of dW needs a bf16 gradient in bf16 mode")
    dW = None
    if need_dW or need_dbias:
        dW = dw_out if (dw_out is not None and dw_out.dtype == torch.float32) else torch.empty(N, d, dtype=torch.float32, device=dev)
        if dW.shape != (N, d):
            raise TypeError("dw_out must be an (N,d) tensor")
    L.call(dev, "rb_ce_bwd", L.ptr(Uc), L.ptr(Wc), L.ptr(b), float(scale), L.ptr(labels.contiguous()), label_base,
                          L.ptr(lse.float().contiguous()), float(grad_scale), L.ptr(grad_scale_dev), M, N, d,
                          L.dtype_code(Uc), mode,
                          L.ptr(dU), L.ptr(dW), L.ptr(db), _count_ptr(n_valid, dev), L.ptr(ws), n, L.stream_ptr(dev))
    return dU, (dW if need_dW else None), db


class _FusedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, U, Wfull, labels, bias, scale, precision, reduction, n_skip, accumulate, n_valid):
        # ``Wfull`` is the whole parameter (N+P, d); the scored table is the view Wfull[n_skip:] (SASRec/main.py:193).
        # Taking the parameter itself lets the backward hand autograd ONE (N+P,d) gradient, written in place by the
        # dW pass, instead of a (N,d) one that the slice's backward would pad into a fresh zero-filled copy.
        W = Wfull[n_skip:] if n_skip else Wfull
        ctx.n_skip = n_skip
        ctx.leaf = Wfull if (accumulate and Wfull.is_leaf and Wfull.requires_grad) else None
        # forward and dU share one sweep whenever the fused pass applies (ctx.needs_input_grad is set in forward)
        ctx.fused_du = bool(ctx.needs_input_grad[0]) and fused_du_supported(U, precision, scale)
        du = m = None
        if ctx.fused_du:
            m, l, ll, du = ce_rowstats(U, W, labels, bias, scale, 0, precision, want_dU=True, n_valid=n_valid)
        else:
            m, l, ll = ce_rowstats(U, W, labels, bias, scale, 0, precision, n_valid=n_valid)
        lse = m + torch.log(l)
        if CHECK_FINITE and not bool(torch.isfinite(lse).all()):   # debugging aid: costs a device->host sync
            raise FloatingPointError("fused_ce: non-finite log-sum-exp (inf/nan logits, or an empty / all -inf row)")
        row_loss = lse - ll
        empty = torch.empty(0, device=U.device)
        ctx.save_for_backward(U, Wfull, labels, bias if bias is not None else empty, lse,
                              du if du is not None else empty, m if du is not None else empty)
        ctx.has_bias = bias is not None
        ctx.scale, ctx.precision, ctx.reduction = scale, precision, reduction
        ctx.n_valid = n_valid
        if n_valid is not None:   # rows beyond the device-side count report lse = label_logit = 0: they add nothing
            total = row_loss.sum()
            return total / n_valid.to(total.dtype).reshape(()) if reduction == "mean" else total
        if reduction == "mean":
            return row_loss.mean()
        if reduction == "sum":
            return row_loss.sum()
        raise ValueError(f"reduction {reduction!r} not supported by the fused path")

    @staticmethod
    def backward(ctx, grad_out):
        U, Wfull, labels, bias, lse, du_un, row_max = ctx.saved_tensors
        bias = bias if ctx.has_bias else None
        P = ctx.n_skip
        W = Wfull[P:] if P else Wfull
        M = U.shape[0]
        nv = ctx.n_valid
        g = 1.0 / (M if (ctx.reduction == "mean" and nv is None) else 1)
        need = ctx.needs_input_grad
        # the upstream scalar stays on the device: no host synchronisation in backward
        gdev = grad_out.detach().float().reshape(1).contiguous()
        if nv is not None and ctx.reduction == "mean":
            gdev = gdev / nv.to(gdev.dtype)   # mean over the device-side row count
        need_db = ctx.has_bias and need[3]
        dU = None
        if ctx.fused_du:
            dU = ce_du_finish(du_un, row_max, lse, W, labels, g, ctx.scale, 0, gdev, n_valid=nv)
        dU2, dW, db = table_gradient(U, Wfull, P, labels, lse, g, bias, ctx.scale, 0, ctx.precision, gdev,
                                     need[0] and not ctx.fused_du, need[1], need_db, ctx.leaf, n_valid=nv)
        dU = dU if dU is not None else dU2
        return (
            dU.to(U.dtype) if dU is not None else None,
            dW,
            None,
            db.to(bias.dtype) if db is not None else None,
            None, None, None, None, None, None,
        )


def table_gradient(U, Wfull, P, labels, lse, g, bias, scale, label_base, precision, gdev, need_du, need_dw, need_db, leaf,
                   n_valid=None):
    """The CE backward's dW / dbias (and dU when the forward did not accumulate it) for a scored table
    ``Wfull[P:]`` -> (dU | None, gradient for ``Wfull`` as autograd wants it | None, dbias | None).

    * ``leaf`` (the parameter itself, accumulate mode) with an allocated bf16 ``.grad``: the dW pass ADDS its rows
      into ``leaf.grad[P:]`` and autograd gets ``None`` for the table; with ``leaf.grad is None`` (bf16) the pass
      writes the new gradient buffer and assigns it;
    * otherwise one (N+P,d) tensor in the parameter's dtype, rows [P:] written in place by the pass, pad rows zero."""
    W = Wfull[P:] if P else Wfull
    bf16_mode = _mode_for(U, precision) == "bf16"
    if (need_dw and not need_du and leaf is not None and leaf.grad is not None and leaf.grad.dtype == torch.bfloat16
            and leaf.grad.is_contiguous() and leaf.grad.shape == Wfull.shape and Wfull.dtype == torch.bfloat16 and bf16_mode):
        try:
            _, _, db = ce_backward(U, W, labels, lse, g, bias, scale, label_base, False, True, need_db, precision,
                                   grad_scale_dev=gdev, dw_dtype=torch.bfloat16, dw_out=leaf.grad[P:], accumulate=True,
                                   n_valid=n_valid)
            return None, None, db
        except NotImplementedError:
            pass   # several splits / very many query rows: a separate dW below, autograd accumulates it
    if (need_dw and not need_du and leaf is not None and leaf.grad is None and Wfull.dtype == torch.bfloat16 and bf16_mode
            and Wfull.is_contiguous()):
        # ``zero_grad()`` left no gradient (set_to_none=True, torch's default): the pass WRITES the parameter's new
        # gradient buffer -- what autograd does with the first gradient of a leaf -- and whoever comes next in this
        # backward (the gather's scatter-add) adds into it.  No zero-fill of the (N+P,d) buffer, no read of old rows.
        buf = torch.empty_like(Wfull)
        if P:
            buf[:P].zero_()
        _, _, db = ce_backward(U, W, labels, lse, g, bias, scale, label_base, False, True, need_db, precision,
                               grad_scale_dev=gdev, dw_dtype=torch.bfloat16, dw_out=buf[P:], n_valid=n_valid)
        leaf.grad = buf
        return None, None, db
    dfull = dw_out = None
    if need_dw and P and (Wfull.dtype == torch.float32 or (Wfull.dtype == torch.bfloat16 and bf16_mode)):
        dfull = torch.empty_like(Wfull)     # the pass writes rows [P:] in place, the P pad rows are zero
        dfull[:P].zero_()
        dw_out = dfull[P:]
    dU, dW, db = ce_backward(U, W, labels, lse, g, bias, scale, label_base, need_du, need_dw, need_db, precision,
                             grad_scale_dev=gdev, dw_dtype=W.dtype, dw_out=dw_out, n_valid=n_valid)
    if dW is not None:
        if dfull is not None:
            dW = dfull
        elif P:   # dtype combinations the in-place route does not cover
            dW = torch.cat([torch.zeros(P, dW.shape[1], dtype=dW.dtype, device=dW.device), dW])
        dW = dW.to(Wfull.dtype)
    return dU, dW, db


def fused_ce(U: torch.Tensor, W: torch.Tensor, labels: torch.Tensor, bias: Optional[torch.Tensor] = None,
             scale: float = 1.0, precision: Optional[str] = None, reduction: str = "mean", n_skip: int = 0,
             accumulate: bool = False, n_valid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Drop-in for ``self.criterion(torch.einsum("MD,ND->MN", U, W), labels)`` with
    ``criterion = CrossEntropy4Logits(reduction="mean")`` (SASRec/main.py:126,217-219): same value,
    same gradients, no (M,N) logit matrix in forward or backward.

    ``n_skip = P``: ``W`` is the whole (N+P, d) embedding parameter and the scored table is ``W[P:]``
    (``self.Item.embeddings.weight[self.NUM_PADS:]``, SASRec/main.py:193; labels index that view).  The gradient
    then comes back as one (N+P,d) tensor written in place -- no zero-padded copy from the slice's backward.

    ``accumulate=True`` (bf16 parameter): the dW pass adds its rows straight into ``W.grad`` -- or, after a
    ``zero_grad()`` that set it to None, writes the new ``W.grad`` -- and together with
    ``gather_rows(..., accumulate=True)`` the table's gradient is built in ONE buffer without any table-sized temporary.  Same caveat as there: not for ``torch.autograd.grad``.

    ``n_valid`` (int32 device scalar, from ``compact_queries``): only the first ``n_valid`` rows of ``U`` / ``labels``
    exist, the rest is capacity; the mean runs over ``n_valid`` and no count ever travels to the host."""
    if U.shape[0] == 0:
        # no query rows (e.g. a BERT4Rec step whose random mask selected nothing): F.cross_entropy returns
        # NaN for "mean" and 0 for "sum", with zero gradients -- reproduced without a launch
        zero = (U.sum() + W.sum() * 0 + (bias.sum() * 0 if bias is not None else 0)).float()
        return zero / 0 if reduction == "mean" else zero
    if n_skip and bias is not None:
        raise ValueError("n_skip is for bias-free embedding tables (a bias head scores every row of its weight)")
    if U.shape[1] % 8:   # e.g. d = 50: zero columns (a padded copy per call; the in-place gradient routes do not apply)
        U, W = _pad_cols(U), _pad_cols(W)
    return _FusedCE.apply(U, W, labels, bias, float(scale), precision, reduction, int(n_skip), bool(accumulate), n_valid)


# --------------------------------------------------------------------------------------
# masked full-catalog top-K
# --------------------------------------------------------------------------------------
def topk_eval(U: torch.Tensor, W: torch.Tensor, K: int, seen_crow: Optional[torch.Tensor] = None,
              seen_col: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None, scale: float = 1.0,
              id_base: int = 0, precision: Optional[str] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-K (score desc, id asc) of ``scale*U W^T + bias`` with the row's seen ids skipped, i.e. the
    outcome of ``scores[seen] = -1e23; torch.topk(scores, K)`` (UniSRec/main.py:408-413 + the metric
    functions) without the dense (B,N).  ``seen_col`` must be sorted ascending inside each row.
    Returns (vals fp32 (B,K), ids int32 (B,K)); missing entries are (-1e23, -1)."""
    dev = L.require_cuda(U, W, seen_crow, seen_col, bias)
    Uc, Wc, mode = _prep(U.detach(), W.detach(), precision)
    B, d = Uc.shape
    N = Wc.shape[0]
    b = None if bias is None else bias.detach().float().contiguous()
    nnz = 0
    if seen_crow is not None:
        if seen_crow.dtype != torch.int64 or seen_col.dtype != torch.int64:
            raise TypeError("seen CSR must be int64")
        if seen_crow.numel() != B + 1:
            raise ValueError("seen_crow must have B+1 entries")
        nnz = seen_col.numel()
        if nnz == 0:  # nothing to mask: same as no seen lists
            seen_crow = seen_col = None
        else:
            seen_crow, seen_col = seen_crow.contiguous(), seen_col.contiguous()
    vals = torch.empty(B, K, dtype=torch.float32, device=dev)
    ids = torch.empty(B, K, dtype=torch.int32, device=dev)
    ws, n = _ws(dev, L.OP_TOPK_EVAL, B, N, d, K=K, mode=mode, nnz=nnz)
    L.call(dev, "rb_topk_eval", L.ptr(Uc), L.ptr(Wc), L.ptr(b), float(scale), L.ptr(seen_crow), L.ptr(seen_col), nnz,
                             id_base, B, N, d, L.dtype_code(Uc), mode, K, L.ptr(vals), L.ptr(ids), L.ptr(ws), n,
                             L.stream_ptr(dev))
    return vals, ids


def topk_merge_packed(packed: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge of per-shard sorted lists in the layout one all-gather leaves them: packed (R,2,B,K) int32, plane 0 the
    float32 values' bits, plane 1 the ids -> the global (B,K) list."""
    dev = L.require_cuda(packed)
    R, two, B, K = packed.shape
    if packed.dtype != torch.int32 or two != 2 or not packed.is_contiguous():
        raise TypeError("packed must be a contiguous int32 (R,2,B,K) tensor")
    ov = torch.empty(B, K, dtype=torch.float32, device=dev)
    oi = torch.empty(B, K, dtype=torch.int32, device=dev)
    L.call(dev, "rb_topk_merge_packed", L.ptr(packed), R, B, K, L.ptr(ov), L.ptr(oi), L.stream_ptr(dev))
    return ov, oi


def topk_merge(vals: torch.Tensor, ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge per-shard sorted lists (R,B,K) into the global (B,K) list."""
    dev = L.require_cuda(vals, ids)
    R, B, K = vals.shape
    ov = torch.empty(B, K, dtype=torch.float32, device=dev)
    oi = torch.empty(B, K, dtype=torch.int32, device=dev)
    L.call(dev, "rb_topk_merge", L.ptr(vals.float().contiguous()), L.ptr(ids.int().contiguous()), R, B, K,
                              L.ptr(ov), L.ptr(oi), L.stream_ptr(dev))
    return ov, oi
