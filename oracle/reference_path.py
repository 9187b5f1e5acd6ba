"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the full-catalog scoring path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product (``recboard_b200/``)
never does: it fails loudly when its CUDA library is missing.

What is restated here, and from where (paths relative to /root/reference):

  * item gather            ``self.Item.embeddings(seqs)``          SASRec/main.py:183
  * its backward           ``embedding_dense_backward`` (padding row zeroed; autograd of :183, run at :249)
  * score contraction      ``einsum("MD,ND->MN")`` / ``("BD,ND->BN")`` / ``("BKD,ND->BN")``
                           SASRec/main.py:217,228; MF-BPR/main.py:104; LightGCN/main.py:120;
                           ``self.fc(userEmbds)`` (+bias) BERT4Rec/main.py:181,189; HSTU ``/ cfg.temperature`` HSTU/main.py:197
  * CE loss                ``CrossEntropy4Logits(reduction="mean")`` = ``F.cross_entropy`` SASRec/main.py:126,219
  * evaluate               line-for-line from UniSRec/main.py:400-447 (mask value -1e23 applied
                           BEFORE ranking, dense multi-hot targets, ``n=bsz`` weighting)
  * pool / sampled scoring ``itemEmbds[ids]`` + ``einsum("MD,MKD->MK")``  SASRec/main.py:230-236, HSTU/main.py:192-197
  * row normalisation      ``F.normalize(weight[NUM_PADS:], dim=-1)``       HSTU/main.py:180-184
  * LightGCN propagation   ``self.Adj @ allEmbds``                            LightGCN/main.py:83
  * HR@k / NDCG@k / RECALL / PRECISION / MRR: ``freerec.metrics`` is NOT in the tree.  The
    formulas below are the standard ones (topk -> gather -> hit / DCG/IDCG); for the LOU
    protocol (exactly one target per row, HSTU/sampler.py:124) they are unambiguous.

PARITY PIN STATUS
  * model half (gather, contraction, CE, gradients, pool / sampled scoring, normalisation,
    propagation): PINNED against outputs of the reference's
    own unmodified code imported in the build container (``oracle/gen_golden.py`` ->
    ``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).
  * evaluate half -- seen masking (``scores[seen] = -1e23`` before ranking) and dense targets: PINNED against the
    tensors the reference's own ``CoachForUniSRec.evaluate`` (UniSRec/main.py:400-447, imported unmodified) hands to
    its metric functions (``gen_golden.gen_unisrec_evaluate`` -> ``tests/golden/unisrec_evaluate.npz``).
  * metric functions (HITRATE / NDCG / RECALL / PRECISION / MRR): they live in un-vendored ``freerec==1.0.1`` and
    cannot be executed here -> "parity unpinned" for the multi-target IDCG convention (the single-target LOU case
    every shipped config uses is convention-free: hit = rank < k, ndcg = 1 / log2(rank + 2)).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

MASK_VALUE = -1e23  # UniSRec/main.py:413


# --------------------------------------------------------------------------- a2/a3
def gather_rows(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """out[..., :] = table[idx[...], :]   (SASRec/main.py:183)."""
    return F.embedding(idx, table)


def scatter_add_rows(
    grad_out: torch.Tensor, idx: torch.Tensor, n_rows: int, padding_idx: int = -1
) -> torch.Tensor:
    """Dense embedding backward: dTable[idx] += dOut; row ``padding_idx`` forced to 0."""
    d = grad_out.shape[-1]
    g = torch.zeros(n_rows, d, dtype=torch.float32)
    flat = idx.reshape(-1)
    go = grad_out.reshape(-1, d).float()
    if padding_idx >= 0:
        keep = flat != padding_idx
        flat, go = flat[keep], go[keep]
    g.index_add_(0, flat, go)
    return g


# ------------------------------------------------------------------------------ a5
def score_dense(
    U: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor] = None, scale: float = 1.0
) -> torch.Tensor:
    """S = scale * (U @ W^T) + bias   (SASRec/main.py:217,228; BERT4Rec/main.py:181).  float32 like the reference;
    float64 inputs stay float64 (the high-precision arbiter of the full-size parity checks)."""
    dt = torch.float64 if U.dtype == torch.float64 else torch.float32
    S = torch.einsum("MD,ND->MN", U.to(dt), W.to(dt))
    if scale != 1.0:
        S = S * scale
    if bias is not None:
        S = S + bias.to(dt)
    return S


# --------------------------------------------------------------------------- a6/a7
def ce_loss(U, W, labels, bias=None, scale: float = 1.0) -> torch.Tensor:
    """mean_i( logsumexp_j S_ij - S_i,label_i )   (SASRec/main.py:217-219)."""
    return F.cross_entropy(score_dense(U, W, bias, scale), labels, reduction="mean")


def gather_dot(U, table, idx, scale: float = 1.0):
    """``einsum("MD,MKD->MK", U, table[idx]) * scale`` -- the pool / sampled scoring lines
    (SASRec/main.py:230-236, HSTU/main.py:192-197, MF-BPR/main.py:84-91,106-109)."""
    return torch.einsum("MD,MKD->MK", U.float(), table.float()[idx]) * scale


def spmm(A, X):
    """``self.Adj @ allEmbds`` (LightGCN/main.py:83) on the host."""
    return torch.sparse.mm(A, X.float()) if A.layout != torch.sparse_csr else A @ X.float()


def normalize_rows(x, eps: float = 1e-12):
    """``F.normalize(x, dim=-1)`` -- the table / user normalisation of HSTU/main.py:180-184."""
    return torch.nn.functional.normalize(x.float(), dim=-1, eps=eps)


def ce_fwd_bwd(U, W, labels, bias=None, scale: float = 1.0, grad_out: float = 1.0):
    """Loss and its gradients w.r.t. U, W (and bias) through autograd -- exactly what
    ``loss.backward()`` (SASRec/main.py:249) produces for the lines :217-219."""
    dt = torch.float64 if U.dtype == torch.float64 else torch.float32
    U = U.detach().to(dt).requires_grad_(True)
    W = W.detach().to(dt).requires_grad_(True)
    b = bias.detach().to(dt).requires_grad_(True) if bias is not None else None
    loss = ce_loss(U, W, labels, b, scale)
    (loss * grad_out).backward()
    return (
        loss.detach(), U.grad, W.grad, (b.grad if b is not None else None)
    )


def ce_fwd_bwd_chunked(U, W, labels, bias=None, scale: float = 1.0, chunk: int = 512):
    """The same loss and gradients as ``ce_fwd_bwd`` from the closed form of the CE gradient
    (G = (softmax(S) - onehot) / M; dU = scale G W; dW = scale G^T U; dbias = sum_i G), query rows ``chunk`` at a time so
    that only a (chunk, N) block of logits exists -- the arbiter of the full-size parity tests (run in float64, on
    whatever device the inputs live on; checked against ``ce_fwd_bwd`` itself in tests/test_oracle_golden.py)."""
    dt = torch.float64 if U.dtype == torch.float64 else torch.float32
    U, W = U.to(dt), W.to(dt)
    b = bias.to(dt) if bias is not None else None
    M = U.shape[0]
    dU = torch.empty_like(U)
    dW = torch.zeros_like(W)
    db = torch.zeros(W.shape[0], dtype=dt, device=W.device) if b is not None else None
    loss = torch.zeros((), dtype=dt, device=U.device)
    for lo in range(0, M, chunk):
        hi = min(lo + chunk, M)
        S = score_dense(U[lo:hi], W, b, scale)
        lse = torch.logsumexp(S, dim=1)
        lab = labels[lo:hi]
        loss += (lse - S.gather(1, lab[:, None]).squeeze(1)).sum()
        G = torch.exp(S - lse[:, None])
        G[torch.arange(hi - lo, device=G.device), lab] -= 1.0
        G /= M
        dU[lo:hi] = scale * (G @ W)
        dW += scale * (G.T @ U[lo:hi])
        if db is not None:
            db += G.sum(0)
    return loss / M, dU, dW, db


def ce_rowstats(U, W, labels, bias=None, scale: float = 1.0):
    """(row_max, row_sumexp, label_logit) -- the partials a shard reports."""
    S = score_dense(U, W, bias, scale)
    m = S.max(dim=1).values
    l = torch.exp(S - m[:, None]).sum(dim=1)
    ll = S.gather(1, labels[:, None]).squeeze(1)
    return m, l, ll


# -------------------------------------------------------------------------- a8-a10
def lists_to_csr(rows: Sequence[Sequence[int]]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Ragged id lists -> (crow[B+1], col[nnz]) sorted-unique per row (``Field.to_csr``)."""
    crow, col = [0], []
    for r in rows:
        r = sorted(set(int(x) for x in r))
        col.extend(r)
        crow.append(len(col))
    return torch.tensor(crow, dtype=torch.int64), torch.tensor(col, dtype=torch.int64)


def csr_to_dense(crow: torch.Tensor, col: torch.Tensor, n_cols: int) -> torch.Tensor:
    B = crow.numel() - 1
    dense = torch.zeros(B, n_cols, dtype=torch.float32, device=crow.device)
    rows = torch.repeat_interleave(torch.arange(B, device=crow.device), crow[1:] - crow[:-1])
    dense[rows, col] = 1.0
    return dense


def mask_seen(scores: torch.Tensor, seen_crow, seen_col) -> torch.Tensor:
    """``scores[seen] = -1e23``   (UniSRec/main.py:409-413)."""
    seen = csr_to_dense(seen_crow, seen_col, scores.shape[1]).bool()
    scores = scores.clone()
    scores[seen] = MASK_VALUE
    return scores


def topk_sorted(scores: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k by (score desc, id asc).  ``torch.topk`` leaves the order among equal scores
    unspecified; the build fixes it to lowest-id-first and so does the oracle."""
    vals, ids = torch.sort(scores, dim=1, descending=True, stable=True)
    return vals[:, :k].contiguous(), ids[:, :k].contiguous()


def _dcg_weights(k: int) -> torch.Tensor:
    return 1.0 / torch.log2(torch.arange(k, dtype=torch.float32) + 2.0)


#: "stable": (score desc, id asc) -- the documented tie policy, used by every parity test.
#: "torch":  plain ``torch.topk`` exactly as the reference's metric functions call it (tie order
#:           unspecified) -- used only when TIMING the reference path on the CPU (bench.py).
TOPK_IMPL = "stable"


def _topk_idx(scores, k):
    if TOPK_IMPL == "torch":
        return torch.topk(scores, k, dim=1).indices
    return topk_sorted(scores, k)[1]


def hitrate(scores, targets, k):
    idx = _topk_idx(scores, k)
    return (targets.gather(1, idx).sum(-1) > 0).float()


def recall(scores, targets, k):
    idx = _topk_idx(scores, k)
    hits = targets.gather(1, idx).sum(-1)
    return hits / targets.sum(-1).clamp_min(1.0)


def precision(scores, targets, k):
    idx = _topk_idx(scores, k)
    return targets.gather(1, idx).sum(-1) / k


def ndcg(scores, targets, k):
    idx = _topk_idx(scores, k)
    h = targets.gather(1, idx)
    w = _dcg_weights(k)
    dcg = (h * w).sum(-1)
    n_rel = targets.sum(-1).clamp(max=k).long()
    idcg = torch.cumsum(w, 0)[(n_rel - 1).clamp_min(0)]
    return torch.where(n_rel > 0, dcg / idcg, torch.zeros_like(dcg))


def mrr(scores, targets, k=None):
    k = scores.shape[1] if k is None else k
    idx = _topk_idx(scores, k)
    h = targets.gather(1, idx)
    first = torch.where(h.sum(-1) > 0, h.argmax(-1), torch.full_like(h[:, 0], -1, dtype=torch.long))
    return torch.where(first >= 0, 1.0 / (first.float() + 1.0), torch.zeros(len(h)))


METRICS = {"HITRATE": hitrate, "NDCG": ndcg, "RECALL": recall, "PRECISION": precision, "MRR": mrr}


class AverageMeter:
    """bsz-weighted running mean -- ``monitor(..., n=bsz, reduction="mean")`` (UniSRec/main.py:428-435)."""

    def __init__(self):
        self.sum, self.n = 0.0, 0

    def update(self, batch_mean: float, n: int):
        self.sum += float(batch_mean) * n
        self.n += n

    @property
    def avg(self) -> float:
        return self.sum / max(self.n, 1)


def evaluate_batch(
    scores: torch.Tensor,
    seen_crow: Optional[torch.Tensor],
    seen_col: Optional[torch.Tensor],
    target_crow: torch.Tensor,
    target_col: torch.Tensor,
    monitors: Sequence[str],
) -> Dict[str, float]:
    """One iteration of ``Coach.evaluate`` (UniSRec/main.py:403-435) on a dense (B,N) score
    matrix: mask seen, build dense targets, one metric call per ``METRIC@k`` monitor,
    batch mean as float32."""
    if seen_crow is not None:
        scores = mask_seen(scores, seen_crow, seen_col)
    targets = csr_to_dense(target_crow, target_col, scores.shape[1])
    out = {}
    for mon in monitors:
        name, k = mon.split("@")
        out[mon.upper()] = METRICS[name.upper()](scores, targets, int(k)).mean().item()
    return out


def evaluate_sweep(batches, monitors) -> Dict[str, float]:
    """Whole evaluation sweep: bsz-weighted mean of per-batch float32 means."""
    meters = {m.upper(): AverageMeter() for m in monitors}
    for scores, seen_crow, seen_col, tcrow, tcol in batches:
        res = evaluate_batch(scores, seen_crow, seen_col, tcrow, tcol, monitors)
        for k, v in res.items():
            meters[k].update(v, scores.shape[0])
    return {k: m.avg for k, m in meters.items()}


# ------------------------------------------------------- shard merge math
def merge_rowstats(parts: List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]):
    """Merge per-shard (max, sumexp, label_logit) into (lse, label_logit)."""
    m = torch.stack([p[0] for p in parts]).max(0).values
    l = sum(p[1] * torch.exp(p[0] - m) for p in parts)
    ll = sum(p[2] for p in parts)
    return m + torch.log(l), ll


def merge_topk(parts: List[Tuple[torch.Tensor, torch.Tensor]], k: int):
    """Merge per-shard sorted (vals, global ids) lists into the global top-k (score desc, id asc)."""
    vals = torch.cat([p[0] for p in parts], dim=1)
    ids = torch.cat([p[1] for p in parts], dim=1).long()
    # sort by id first (stable), then by value desc (stable) => (val desc, id asc)
    o = torch.argsort(ids, dim=1, stable=True)
    vals, ids = vals.gather(1, o), ids.gather(1, o)
    o = torch.argsort(vals, dim=1, descending=True, stable=True)
    return vals.gather(1, o)[:, :k], ids.gather(1, o)[:, :k]


def hits_from_topk(top_ids: torch.Tensor, target_crow, target_col, n_items: int) -> torch.Tensor:
    """(B,K) float32 hit matrix of a ranked id list against the dense multi-hot targets of
    UniSRec/main.py:414 (``targets.gather(1, idx)``); missing entries (id < 0) never hit."""
    targets = csr_to_dense(target_crow, target_col, n_items)
    return targets.gather(1, top_ids.long().clamp_min(0)) * (top_ids >= 0).float()


def metrics_from_topk(top_ids: torch.Tensor, target_crow, target_col, n_items: int, monitors):
    """Same metric values computed from a (B,Kmax) id list instead of dense scores
    (the fused path's route; must agree with ``evaluate_batch`` whenever ranks agree)."""
    targets = csr_to_dense(target_crow, target_col, n_items)
    ids = top_ids.long().clamp_min(0)
    valid = (top_ids >= 0).float()
    h_all = targets.gather(1, ids) * valid
    out = {}
    for mon in monitors:
        name, k = mon.split("@")
        name, k = name.upper(), int(k)
        h = h_all[:, :k]
        if name == "HITRATE":
            v = (h.sum(-1) > 0).float()
        elif name == "RECALL":
            v = h.sum(-1) / targets.sum(-1).clamp_min(1.0)
        elif name == "PRECISION":
            v = h.sum(-1) / k
        elif name == "NDCG":
            w = _dcg_weights(k)
            dcg = (h * w).sum(-1)
            n_rel = targets.sum(-1).clamp(max=k).long()
            idcg = torch.cumsum(w, 0)[(n_rel - 1).clamp_min(0)]
            v = torch.where(n_rel > 0, dcg / idcg, torch.zeros_like(dcg))
        elif name == "MRR":
            first = torch.where(h.sum(-1) > 0, h.argmax(-1), torch.full_like(h[:, 0], -1, dtype=torch.long))
            v = torch.where(first >= 0, 1.0 / (first.float() + 1.0), torch.zeros(len(h)))
        else:
            raise KeyError(name)
        out[mon.upper()] = v.mean().item()
    return out
