"""HR@k / NDCG@k / RECALL@k / PRECISION@k / MRR@k from a sorted top-K id list.

The reference computes these inside ``Coach.monitor(scores, targets, ...)`` from dense (B,N) score
and multi-hot target matrices, one ``torch.topk`` per ``metric@k`` (UniSRec/main.py:428-435; the
metric functions are freerec's).  The fused path gets the sorted (B,Kmax) id list from
``ops.topk_eval`` once and derives every metric@k from it.

Reduction contract: per batch a float32 mean, then a bsz-weighted running mean
(``monitor(..., n=bsz, reduction="mean")``).  ``hits_from_topk`` is a kernel (``rb_topk_hits``; CUDA tensors
only); the tiny (B,K) hit matrix is reduced with the same float32 torch ops the CPU oracle uses, so metric
values are bit-identical whenever the ranked ids agree.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch


def hits_from_topk(top_ids: torch.Tensor, target_crow: torch.Tensor, target_col: torch.Tensor,
                   n_items: int) -> torch.Tensor:
    """(B,K) float32 hit matrix: 1 where the ranked id is one of the row's targets.

    ``target_crow/col`` is the CSR of ``data[IUnseen]`` (``Item.to_csr``, UniSRec/main.py:414);
    LOU evaluation has exactly one target per row."""
    from . import _lib as L
    B, K = top_ids.shape
    dev = L.require_cuda(top_ids, target_crow, target_col)   # raises on CPU tensors: the product has no CPU path
    ids32 = top_ids.to(torch.int32).contiguous()
    hits = torch.empty(B, K, dtype=torch.float32, device=dev)
    L.call(dev, "rb_topk_hits", L.ptr(ids32), L.ptr(target_crow.to(torch.int64).contiguous()),
           L.ptr(target_col.to(torch.int64).contiguous()), B, K, L.ptr(hits), L.stream_ptr(dev))
    return hits


def _dcg_weights(k: int) -> torch.Tensor:
    return 1.0 / torch.log2(torch.arange(k, dtype=torch.float32) + 2.0)


def metric_rows(hits: torch.Tensor, n_targets: torch.Tensor, name: str, k: int) -> torch.Tensor:
    """Per-row metric values (float32, on ``hits.device``)."""
    h = hits[:, :k]
    if name == "HITRATE":
        return (h.sum(-1) > 0).float()
    if name == "RECALL":
        return h.sum(-1) / n_targets.clamp_min(1.0)
    if name == "PRECISION":
        return h.sum(-1) / k
    if name == "NDCG":
        w = _dcg_weights(k).to(h.device)
        dcg = (h * w).sum(-1)
        n_rel = n_targets.clamp(max=k).long()
        idcg = torch.cumsum(w, 0)[(n_rel - 1).clamp_min(0)]
        return torch.where(n_rel > 0, dcg / idcg, torch.zeros_like(dcg))
    if name == "MRR":
        first = torch.where(h.sum(-1) > 0, h.argmax(-1), torch.full_like(h[:, 0], -1, dtype=torch.long))
        return torch.where(first >= 0, 1.0 / (first.float() + 1.0), torch.zeros(len(h), device=h.device))
    raise KeyError(f"unknown metric {name!r}")


_KINDS = {"HITRATE": 0, "RECALL": 1, "PRECISION": 2, "NDCG": 3, "MRR": 4}
_W_CACHE: Dict = {}


def batch_metrics_device(top_ids: torch.Tensor, target_crow: torch.Tensor, target_col: torch.Tensor,
                         monitors: Sequence[str]) -> torch.Tensor:
    """float32 device vector of the batch means of ``monitors`` (same order), computed from the ranked ids in ONE pass
    (``rb_topk_metrics``): no hit matrix, no per-metric launches, nothing read back -- the caller decides when the
    ``len(monitors)`` floats travel to the host (e.g. once per sweep, or overlapped with the next batch)."""
    import ctypes as C
    from . import _lib as L
    B, K = top_ids.shape
    dev = L.require_cuda(top_ids, target_crow, target_col)
    kinds, ks = [], []
    for mon in monitors:
        name, k = mon.split("@")
        if name.upper() not in _KINDS:
            raise KeyError(f"unknown metric {name!r}")
        if int(k) > K:
            raise ValueError(f"{mon}: k exceeds the ranked list length {K}")
        kinds.append(_KINDS[name.upper()]); ks.append(int(k))
    key = (dev, K)
    if key not in _W_CACHE:   # the oracle's float32 discount weights and their running sum
        w = _dcg_weights(K)
        _W_CACHE[key] = (w.to(dev), torch.cumsum(w, 0).to(dev))
    w, w_cum = _W_CACHE[key]
    n = len(kinds)
    blocks = L.lib().rb_topk_metrics_blocks(B)
    partial = torch.empty(blocks * 32, dtype=torch.float64, device=dev)
    out = torch.empty(n, dtype=torch.float32, device=dev)
    L.call(dev, "rb_topk_metrics", L.ptr(top_ids.to(torch.int32).contiguous()), L.ptr(target_crow.to(torch.int64).contiguous()),
           L.ptr(target_col.to(torch.int64).contiguous()), B, K, L.ptr(w), L.ptr(w_cum), (C.c_int32 * n)(*kinds),
           (C.c_int32 * n)(*ks), n, L.ptr(partial), L.ptr(out), L.stream_ptr(dev))
    return out


def batch_metrics(top_ids: torch.Tensor, target_crow: torch.Tensor, target_col: torch.Tensor, n_items: int,
                  monitors: Sequence[str], exact: bool = True) -> Dict[str, float]:
    """Batch means for every ``METRIC@k`` in ``monitors`` (names as in ``cfg.monitors``,
    e.g. SASRec/configs/Amazon2014Beauty_550_LOU.yaml:21).

    exact=True reduces the (B,K) hit matrix with CPU float32 torch ops (bit-identical to the
    oracle); exact=False computes all of them on the device in one pass over the ranked ids (``batch_metrics_device``:
    per-row values in float32, batch sums in double -- within 1e-6 of the exact mode) and reads ONE small vector back."""
    if not exact:
        vals = batch_metrics_device(top_ids, target_crow, target_col, monitors).cpu()
        return {mon.upper(): float(v) for mon, v in zip(monitors, vals)}
    hits = hits_from_topk(top_ids, target_crow, target_col, n_items)
    n_t = (target_crow[1:] - target_crow[:-1]).float()
    hits, n_t = hits.cpu(), n_t.cpu()
    return metrics_from_hits(hits, n_t, monitors)


def metrics_from_hits(hits: torch.Tensor, n_t: torch.Tensor, monitors: Sequence[str]) -> Dict[str, float]:
    """Batch means of every ``METRIC@k`` from a (B,K) hit matrix and the per-row target counts (same float32
    torch reductions as the oracle when both live on the host)."""
    out = {}
    for mon in monitors:
        name, k = mon.split("@")
        k = int(k)
        if k > hits.shape[1]:
            raise ValueError(f"{mon}: k exceeds the ranked list length {hits.shape[1]}")
        out[mon.upper()] = metric_rows(hits, n_t, name.upper(), k).mean().item()
    return out


class AverageMeter:
    """bsz-weighted running mean (``monitor(..., n=bsz, reduction="mean")``)."""

    def __init__(self):
        self.sum, self.n = 0.0, 0

    def update(self, batch_mean: float, n: int):
        self.sum += float(batch_mean) * n
        self.n += n

    @property
    def avg(self) -> float:
        return self.sum / max(self.n, 1)


def kmax_of(monitors: Sequence[str]) -> int:
    return max(int(m.split("@")[1]) for m in monitors if "@" in m)


def lists_to_csr(rows, device: Optional[torch.device] = None):
    """Ragged id lists -> (crow, col) int64 with ids sorted-unique per row (``Field.to_csr``)."""
    crow, col = [0], []
    for r in rows:
        r = sorted(set(int(x) for x in (r.tolist() if isinstance(r, torch.Tensor) else r)))
        col.extend(r)
        crow.append(len(col))
    crow = torch.tensor(crow, dtype=torch.int64, device=device)
    col = torch.tensor(col, dtype=torch.int64, device=device)
    return crow, col
