"""Fused evaluation sweep: the override point of ``freerec.launcher.Coach.evaluate`` whose in-tree
statement is UniSRec/main.py:400-447.

Reference, per batch:   scores = model(data, ranking="full")            dense (B,N)         :408
                        seen = Item.to_csr(data[ISeen]).to_dense().bool(); scores[seen] = -1e23  :410-413
                        targets = Item.to_csr(data[IUnseen]).to_dense()  dense (B,N)         :414
                        monitor(scores, targets, n=bsz, ...)  -> one torch.topk per metric@k :428-435

Fused, per batch:       ids = model.recommend_topk(data, Kmax, seen CSR)  one kernel sweep, (B,Kmax)
                        metric@k for every k from that one sorted list (recboard_b200.metrics)
                        monitor(precomputed batch means, n=bsz) through identity metrics
                        (the pattern of TIGER/train_rqvae.py:224-230)

``FusedEvalCoach`` is a mixin for a freerec ``Coach`` subclass; ``evaluate_sweep`` is the same loop
without freerec (used by bench.py and the tests).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence, Tuple

import torch

from . import metrics as MX


def _to_csr(field_value, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """``data[ISeen]`` / ``data[IUnseen]`` arrive as ragged lists or (B,k) tensors of 0-based item ids."""
    if isinstance(field_value, tuple) and len(field_value) == 2 and all(isinstance(x, torch.Tensor) for x in field_value):
        crow, col = field_value
        return crow.to(device), col.to(device)
    if isinstance(field_value, torch.Tensor) and field_value.dim() == 2:
        B, k = field_value.shape
        vals, _ = torch.sort(field_value.to(device), dim=1)
        if k > 1:  # unique per row
            keep = torch.ones_like(vals, dtype=torch.bool)
            keep[:, 1:] = vals[:, 1:] != vals[:, :-1]
            crow = torch.zeros(B + 1, dtype=torch.int64, device=device)
            crow[1:] = keep.sum(1).cumsum(0)
            return crow, vals[keep].contiguous()
        return torch.arange(B + 1, device=device, dtype=torch.int64), vals.reshape(-1).contiguous()
    return MX.lists_to_csr(field_value, device)


@torch.no_grad()
def evaluate_sweep(model, batches: Iterable[Dict], monitors: Sequence[str], n_items: int, remove_seen: bool = True,
                   seen_key=None, unseen_key=None, size_key=None, exact: bool = True) -> Dict[str, float]:
    """bsz-weighted means of every ``METRIC@k`` over an evaluation sweep (UniSRec/main.py:400-435).
    ``model`` provides ``reset_ranking_buffers()`` and ``recommend_topk(data, K, seen_crow, seen_col)``."""
    seen_key = seen_key if seen_key is not None else model.ISeen
    unseen_key = unseen_key if unseen_key is not None else model.IUnseen
    kmax = MX.kmax_of(monitors)
    meters = {m.upper(): MX.AverageMeter() for m in monitors}
    model.reset_ranking_buffers()                                                   # :401
    for data in batches:
        device = next(model.parameters()).device
        seen_crow = seen_col = None
        if remove_seen:                                                             # :409
            seen_crow, seen_col = _to_csr(data[seen_key], device)
        tcrow, tcol = _to_csr(data[unseen_key], device)                             # :414
        bsz = int(data[size_key]) if size_key is not None and size_key in data else tcrow.numel() - 1   # :403
        _, ids = model.recommend_topk(data, kmax, seen_crow, seen_col)              # :408-413 fused
        for name, v in MX.batch_metrics(ids, tcrow, tcol, n_items, monitors, exact=exact).items():
            meters[name].update(v, bsz)                                             # :428-435
    return {k: m.avg for k, m in meters.items()}


class DeviceEvalSplit:
    """Device-resident seen / target CSR of a whole evaluation split, built ONCE.

    The reference rebuilds the per-row seen and target lists in Python for every batch of every evaluation
    sweep (row format ``{User, ISeq, IUnseen, ISeen}``, HSTU/sampler.py:107-125; ``Item.to_csr(...)`` per
    batch, UniSRec/main.py:410-414).  Here the ragged lists of all evaluation rows are turned into two CSRs on
    the device when the split is set up; a batch is a contiguous row range whose CSR is two slices and one
    subtraction -- no host work, no host->device copy of id lists during the sweep."""

    def __init__(self, seen_rows, target_rows, device):
        self.seen_crow, self.seen_col = MX.lists_to_csr(seen_rows, device)
        self.tgt_crow, self.tgt_col = MX.lists_to_csr(target_rows, device)
        if self.seen_crow.numel() != self.tgt_crow.numel():
            raise ValueError("seen and target lists must describe the same rows")
        self.n_rows = self.seen_crow.numel() - 1
        # row offsets are read on the host when a batch is cut (two ints per CSR per batch)
        self._seen_crow_host = self.seen_crow.cpu()
        self._tgt_crow_host = self.tgt_crow.cpu()

    @classmethod
    def from_rows(cls, rows, user_key, seq_key, unseen_key, seen_key, device, maxlen: Optional[int] = None,
                  num_pads: int = 1, padding_value: int = 0) -> "DeviceEvalSplit":
        """The DataLoader-facing half: consume ONCE the rows the reference's evaluation samplers yield
        -- ``{User: user, ISeq: seq, IUnseen: (positive, ...), ISeen: seen}``, HSTU/sampler.py:107-125
        (``_nextitem_from_full``) -- and keep on the device what every later sweep needs: the seen / target CSRs, the
        user ids, and (with ``maxlen``) the input sequences exactly as the reference's pipes finish them --
        truncated to the last ``maxlen`` items, ids shifted by ``NUM_PADS`` (``add_``) and left-padded with
        ``PADDING_VALUE`` (``lpad_``), SASRec/main.py:150-154.  No DataLoader workers, no per-batch Python after this."""
        users, seen_rows, tgt_rows, seqs = [], [], [], []
        for row in rows:
            users.append(int(row[user_key]))
            seen_rows.append(row[seen_key])
            tgt_rows.append(row[unseen_key])
            if maxlen is not None:
                seq = [int(x) + num_pads for x in row[seq_key]][-maxlen:]
                seqs.append([padding_value] * (maxlen - len(seq)) + seq)
        split = cls(seen_rows, tgt_rows, device)
        split.users = torch.tensor(users, dtype=torch.int64, device=device).unsqueeze(1)        # (R, 1) like data[User]
        split.seqs = torch.tensor(seqs, dtype=torch.int64, device=device) if maxlen is not None else None
        return split

    def data(self, lo: int, hi: int, model) -> Dict:
        """The ``data`` dict of rows [lo, hi) in the reference's batch format (views of the device-resident split)."""
        d = {model.User: self.users[lo:hi]}
        if getattr(self, "seqs", None) is not None:
            d[model.ISeq] = self.seqs[lo:hi]
        return d

    @staticmethod
    def _cut(crow, crow_host, col, lo, hi):
        a, b = int(crow_host[lo]), int(crow_host[hi])
        return (crow[lo:hi + 1] - a).contiguous(), col[a:b].contiguous()

    def batch(self, lo: int, hi: int):
        """-> (seen_crow, seen_col, target_crow, target_col) of rows [lo, hi), row offsets rebased to 0."""
        if not (0 <= lo <= hi <= self.n_rows):
            raise IndexError(f"rows [{lo}, {hi}) outside the split of {self.n_rows} rows")
        return (*self._cut(self.seen_crow, self._seen_crow_host, self.seen_col, lo, hi),
                *self._cut(self.tgt_crow, self._tgt_crow_host, self.tgt_col, lo, hi))


@torch.no_grad()
def evaluate_split(score_topk, split: DeviceEvalSplit, monitors: Sequence[str], n_items: int, batch_size: int,
                   remove_seen: bool = True, exact: bool = True) -> Dict[str, float]:
    """The sweep of ``evaluate_sweep`` over a ``DeviceEvalSplit``: ``score_topk(lo, hi, K, seen_crow, seen_col)``
    returns the sorted (vals, ids) of rows [lo, hi) (e.g. a closure over ``model.recommend_topk``)."""
    kmax = MX.kmax_of(monitors)
    meters = {m.upper(): MX.AverageMeter() for m in monitors}
    acc = None   # exact=False: bsz-weighted sums stay on the device, ONE read-back at the end of the sweep
    for lo in range(0, split.n_rows, batch_size):
        hi = min(lo + batch_size, split.n_rows)
        s_crow, s_col, t_crow, t_col = split.batch(lo, hi)
        _, ids = score_topk(lo, hi, kmax, s_crow if remove_seen else None, s_col if remove_seen else None)
        if exact:
            for name, v in MX.batch_metrics(ids, t_crow, t_col, n_items, monitors, exact=True).items():
                meters[name].update(v, hi - lo)                                     # UniSRec/main.py:428-435, n=bsz
        else:
            v = MX.batch_metrics_device(ids, t_crow, t_col, monitors).double() * (hi - lo)
            acc = v if acc is None else acc + v
    if not exact and acc is not None:
        return {m.upper(): float(x) for m, x in zip(monitors, (acc / split.n_rows).cpu())}
    return {k: m.avg for k, m in meters.items()}


@torch.no_grad()
def evaluate_split_model(model, split: DeviceEvalSplit, monitors: Sequence[str], batch_size: int,
                         remove_seen: bool = True, exact: bool = True) -> Dict[str, float]:
    """A whole evaluation sweep of ``model`` (any fused mixin of ``arch.py``) over a ``DeviceEvalSplit.from_rows`` split:
    ``reset_ranking_buffers()`` once (UniSRec/main.py:401), then contiguous row ranges of the device-resident split."""
    model.reset_ranking_buffers()
    return evaluate_split(lambda lo, hi, k, crow, col: model.recommend_topk(split.data(lo, hi, model), k, crow, col),
                          split, monitors, model.Item.count, batch_size, remove_seen=remove_seen, exact=exact)


class FusedEvalCoach:
    """Mixin for a ``freerec.launcher.Coach`` subclass (listed before ``Coach`` in the bases).

    ``set_other`` registers identity metrics named like the configured monitors so that the
    precomputed batch means flow through ``self.monitor`` unchanged (bsz weighting, best-epoch
    tracking, logging stay freerec's)."""

    def set_other(self):
        for monitor in self.cfg.monitors:
            if "@" not in monitor:
                continue
            metric, k = monitor.split("@")
            self.register_metric(f"{metric.upper()}@{k}", func=lambda x: x, fmt=".4f", best_caster=max)

    @torch.no_grad()
    def evaluate(self, epoch: int, step: int = -1, mode: str = "valid"):
        model = self.get_res_sys_arch()
        model.reset_ranking_buffers()                                               # UniSRec/main.py:401
        monitors = [m for m in self.cfg.monitors if "@" in m]
        kmax = MX.kmax_of(monitors)
        n_items = model.Item.count
        for data in self.dataloader:
            bsz = data[self.Size]                                                   # :403
            data = self.dict_to_device(data)                                        # :405
            if self.cfg.ranking != "full":
                raise NotImplementedError("the fused evaluate covers ranking='full' (UniSRec/main.py:407-414)")
            seen_crow = seen_col = None
            if self.remove_seen:                                                    # :409
                seen_crow, seen_col = _to_csr(data[self.ISeen], self.device)
            tcrow, tcol = _to_csr(data[self.IUnseen], self.device)                  # :414
            _, ids = model.recommend_topk(data, kmax, seen_crow, seen_col)
            for name, v in MX.batch_metrics(ids, tcrow, tcol, n_items, monitors).items():
                self.monitor(v, n=bsz, reduction="mean", mode=mode, pool=[name])    # :428-435
