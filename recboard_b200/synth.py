"""Seeded synthetic inputs of the shapes BASELINE.json names: embeddings ~ N(0,1)/sqrt(d)
scaled so logits have unit-ish variance, Zipf(1.0) item popularity for labels / sequences / seen
lists, log-normal seen-list lengths, one held-out target per row (LOU protocol)."""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch


def zipf_ids(n: int, n_items: int, gen: torch.Generator, device, alpha: float = 1.0) -> torch.Tensor:
    """n item ids in [0, n_items) with P(rank r) ~ 1/(r+1)^alpha, via the inverse CDF of the
    continuous approximation (exact enough for a popularity-skew stress)."""
    u = torch.rand(n, generator=gen, device=device, dtype=torch.float64)
    if alpha == 1.0:
        r = torch.exp(u * math.log(n_items + 1.0)) - 1.0
    else:
        a = 1.0 - alpha
        r = ((u * ((n_items + 1.0) ** a - 1.0)) + 1.0) ** (1.0 / a) - 1.0
    ids = r.long().clamp_(0, n_items - 1)
    # decorrelate popularity rank from row index (hot rows spread over the table)
    return (ids * 2654435761) % n_items


def embeddings(rows: int, d: int, gen: torch.Generator, device, dtype=torch.float32, gain: float = 1.0) -> torch.Tensor:
    x = torch.randn(rows, d, generator=gen, device=device, dtype=torch.float32) * (gain / d ** 0.25)
    return x.to(dtype)


def seen_csr(n_rows: int, n_items: int, gen: torch.Generator, device, mean_len: float = 25.0, sigma: float = 0.8,
             max_len: int = 2000) -> Tuple[torch.Tensor, torch.Tensor]:
    """CSR (crow, col) of sorted-unique seen item ids per row; lengths clamp(LogNormal, 3, max_len)."""
    max_len = min(max_len, n_items // 2)
    lens = torch.exp(torch.randn(n_rows, generator=gen, device=device) * sigma + math.log(mean_len))
    lens = lens.long().clamp_(3, max_len)
    total = int(lens.sum())
    rows = torch.repeat_interleave(torch.arange(n_rows, device=device), lens)
    cols = zipf_ids(total, n_items, gen, device)
    keys = torch.unique(rows * n_items + cols)  # sorted, duplicates inside a row dropped
    rows, cols = keys // n_items, keys % n_items
    crow = torch.zeros(n_rows + 1, dtype=torch.int64, device=device)
    crow[1:] = torch.bincount(rows, minlength=n_rows).cumsum(0)
    return crow, cols.contiguous()


def targets(n_rows: int, n_items: int, gen: torch.Generator, device, seen: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
            frac_in_seen: float = 0.05) -> torch.Tensor:
    """One target id per row; ``frac_in_seen`` of the rows take a target from their own seen list
    (the reference masks it anyway, UniSRec/main.py:413 -> a guaranteed miss)."""
    t = zipf_ids(n_rows, n_items, gen, device)
    if seen is not None and frac_in_seen > 0:
        crow, col = seen
        pick = torch.rand(n_rows, generator=gen, device=device) < frac_in_seen
        has = (crow[1:] - crow[:-1]) > 0
        first = col[crow[:-1].clamp_max(max(col.numel() - 1, 0))]
        t = torch.where(pick & has, first, t)
    return t


def sequences(n_rows: int, maxlen: int, n_items: int, gen: torch.Generator, device, mean_len: float = 5.9,
              num_pads: int = 1) -> torch.Tensor:
    """Left-padded (lpad_, SASRec/main.py:150-154) id sequences, ids offset by NUM_PADS, 0 = pad."""
    lens = torch.poisson(torch.full((n_rows,), mean_len, device=device), generator=gen).long().clamp_(1, maxlen)
    pos = torch.arange(maxlen, device=device).unsqueeze(0)
    valid = pos >= (maxlen - lens).unsqueeze(1)
    ids = zipf_ids(n_rows * maxlen, n_items, gen, device).view(n_rows, maxlen) + num_pads
    return torch.where(valid, ids, torch.zeros_like(ids))
