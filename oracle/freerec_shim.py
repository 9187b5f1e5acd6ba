"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (recboard_b200/).

A minimal stand-in for the third-party ``freerec`` package (pinned only by
``freerec.declare(version="1.0.1")``, /root/reference/SASRec/main.py:7; it is NOT
vendored in /root/reference and NOT installed in this image).  It exists so that
the reference model files can be imported *unmodified* in this container to
(a) validate the hand restatement in ``oracle/reference_path.py`` and
(b) generate the golden vectors committed under ``tests/golden/``
    (see ``oracle/gen_golden.py``).

It provides exactly the surface the six hot-path scripts touch at import time and
inside ``encode / fit / recommend_from_full``:

  freerec.declare, freerec.parser.Parser, freerec.models.{RecSysArch,SeqRecArch,
  GenRecArch}, freerec.criterions.{CrossEntropy4Logits,BCELoss4Logits,BPRLoss},
  freerec.launcher.Coach, freerec.data.{datasets,fields,tags,postprocessing},
  freerec.utils.timemeter, torchdata.datapipes (HSTU/sampler.py:5-11).

Semantics restated (no freerec source is available, so these follow the call
shapes visible in the reference tree):
  * ``RecSysArch.forward(data, ranking=...)`` dispatches to ``fit`` when
    ``self.training`` else ``recommend_from_full`` / ``recommend_from_pool``
    (call shapes: UniSRec/main.py:408,416).
  * ``CrossEntropy4Logits(reduction=)`` == ``F.cross_entropy(logits, labels)``
    (constructed SASRec/main.py:126, called :219).
  * ``BPRLoss`` == mean softplus(neg - pos); ``BCELoss4Logits`` ==
    ``binary_cross_entropy_with_logits`` (call shapes SASRec/main.py:211-215).
"""
from __future__ import annotations

import argparse
import importlib.util
import sys
import types
from pathlib import Path
from typing import Dict, Iterable, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = Path("/root/reference")


# --------------------------------------------------------------------------- tags
class _Tag(str):
    pass


USER, ITEM, ID, SEQUENCE, TIMESTAMP, POSITIVE, NEGATIVE, UNSEEN, SEEN, LABEL, SIZE = (
    _Tag(n)
    for n in (
        "USER", "ITEM", "ID", "SEQUENCE", "TIMESTAMP", "POSITIVE", "NEGATIVE",
        "UNSEEN", "SEEN", "LABEL", "SIZE",
    )
)


# ------------------------------------------------------------------------- fields
class Field(nn.Module):
    """Hashable-by-name nn.Module (models call ``self.Item.add_module("embeddings", ...)``,
    SASRec/main.py:70-77) carrying ``count`` and ``fork``/``to_csr``."""

    def __init__(self, name: str, count: Optional[int] = None, tags: Iterable[str] = ()):
        super().__init__()
        self.name = name
        self.count = count
        self.tags = tuple(tags)

    def fork(self, *tags) -> "Field":
        return Field(self.name, self.count, self.tags + tuple(tags))

    def __hash__(self):
        return hash((self.name, self.tags))

    def __eq__(self, other):
        return isinstance(other, Field) and (self.name, self.tags) == (other.name, other.tags)

    def to_csr(self, rows) -> torch.Tensor:
        """Ragged list-of-lists (or (B,k) tensor) of item ids -> sparse CSR (B, count).

        Call shape: ``Item.to_csr(data[self.ISeen]).to(device).to_dense().bool()``
        (UniSRec/main.py:410-412)."""
        if isinstance(rows, torch.Tensor):
            rows = rows.tolist()
        crow = [0]
        col = []
        for r in rows:
            r = sorted(set(int(x) for x in r))
            col.extend(r)
            crow.append(len(col))
        return torch.sparse_csr_tensor(
            torch.tensor(crow, dtype=torch.long),
            torch.tensor(col, dtype=torch.long),
            torch.ones(len(col), dtype=torch.float32),
            size=(len(rows), self.count),
        )


class _FieldTable(dict):
    def __getitem__(self, key):
        if isinstance(key, tuple):
            key = key[0]
        return super().__getitem__(key)


# ----------------------------------------------------------------------- datasets
class _Split:
    def __init__(self, ds):
        self.ds = ds

    def to_normalized_adj(self, normalization="sym"):
        """Synthetic symmetric-normalised bipartite adjacency (LightGCN/main.py:47-49)."""
        U, N = self.ds.n_users, self.ds.n_items
        g = torch.Generator().manual_seed(7)
        nnz = max(4 * U, 8)
        u = torch.randint(0, U, (nnz,), generator=g)
        i = torch.randint(0, N, (nnz,), generator=g) + U
        idx = torch.stack([torch.cat([u, i]), torch.cat([i, u])])
        A = torch.sparse_coo_tensor(idx, torch.ones(idx.size(1)), (U + N, U + N)).coalesce()
        deg = torch.sparse.sum(A, dim=1).to_dense().clamp_min(1.0)
        r, c = A.indices()
        v = A.values() / (deg[r].sqrt() * deg[c].sqrt())
        return torch.sparse_coo_tensor(A.indices(), v, A.shape).coalesce().to_sparse_csr()

    def __getattr__(self, name):  # datapipe verbs are never executed by the oracle
        raise AttributeError(f"datapipe verb {name!r} is out of scope for the oracle shim")


class RecDataSet:
    def __init__(self, root=None, name=None, tasktag=None, n_users=64, n_items=1000):
        self.n_users, self.n_items = n_users, n_items
        self.fields = _FieldTable(
            {
                USER: Field("USER", n_users, (USER, ID)),
                ITEM: Field("ITEM", n_items, (ITEM, ID)),
                TIMESTAMP: Field("TIMESTAMP", None, (TIMESTAMP,)),
            }
        )

    def train(self):
        return _Split(self)

    valid = test = train


# ------------------------------------------------------------------------- parser
class Parser:
    def __init__(self):
        self._p = argparse.ArgumentParser()
        self._defaults = dict(ranking="full", tasktag=None, device="cpu", retain_seen=False)
        self._overrides: Dict[str, object] = {}

    def add_argument(self, *a, **k):
        self._p.add_argument(*a, **k)

    def set_defaults(self, **k):
        self._defaults.update(k)

    def compile(self):
        ns, _ = self._p.parse_known_args([])
        for k, v in self._defaults.items():
            setattr(self, k, v)
        for k, v in vars(ns).items():
            setattr(self, k, v)
        for k, v in _PENDING_OVERRIDES.items():
            setattr(self, k, v)


_PENDING_OVERRIDES: Dict[str, object] = {}


# ------------------------------------------------------------------------- models
class RecSysArch(nn.Module):
    NUM_PADS = 0
    PADDING_VALUE = 0

    def __init__(self, dataset: RecDataSet):
        super().__init__()
        self.dataset = dataset
        self.fields = dataset.fields
        self.User = dataset.fields[USER]
        self.Item = dataset.fields[ITEM]
        self.ISeq = self.Item.fork(SEQUENCE)
        self.IPos = self.Item.fork(POSITIVE)
        self.INeg = self.Item.fork(NEGATIVE)
        self.IUnseen = self.Item.fork(UNSEEN)
        self.ISeen = self.Item.fork(SEEN)
        self.Label = Field("LABEL", None, (LABEL,))
        self.Size = Field("SIZE", None, (SIZE,))

    @property
    def device(self):
        return next(self.parameters()).device

    def reset_ranking_buffers(self):
        pass

    def forward(self, data, ranking: str = "full"):
        if self.training:
            return self.fit(data)
        if ranking == "full":
            return self.recommend_from_full(data)
        if ranking == "pool":
            return self.recommend_from_pool(data)
        raise NotImplementedError(ranking)


class GenRecArch(RecSysArch):
    pass


class SeqRecArch(RecSysArch):
    NUM_PADS = 1
    PADDING_VALUE = 0


class PredRecArch(RecSysArch):
    pass


# --------------------------------------------------------------------- criterions
class _Crit(nn.Module):
    def __init__(self, reduction: str = "mean"):
        super().__init__()
        self.reduction = reduction

    def regularize(self, params, rtype: str = "l2"):
        params = [params] if isinstance(params, torch.Tensor) else list(params)
        if rtype == "l2":
            return sum(p.pow(2).sum() for p in params) / 2
        if rtype == "l1":
            return sum(p.abs().sum() for p in params)
        raise NotImplementedError(rtype)


class CrossEntropy4Logits(_Crit):
    def forward(self, logits, targets):
        return F.cross_entropy(logits, targets, reduction=self.reduction)


class BCELoss4Logits(_Crit):
    def forward(self, logits, targets):
        return F.binary_cross_entropy_with_logits(logits, targets, reduction=self.reduction)


class BPRLoss(_Crit):
    def forward(self, pos, neg):
        loss = F.softplus(neg - pos)
        return loss.mean() if self.reduction == "mean" else loss.sum()


class Coach:
    def __init__(self, *a, **k):
        pass


# ------------------------------------------------------------------ installation
def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install() -> types.ModuleType:
    """Insert the stub ``freerec`` (and ``torchdata.datapipes``) into ``sys.modules``."""
    if "freerec" in sys.modules and getattr(sys.modules["freerec"], "_IS_SHIM", False):
        return sys.modules["freerec"]

    def timemeter(f=None, *a, **k):
        if callable(f):
            return f
        return lambda g: g

    fr = _mod("freerec", _IS_SHIM=True, declare=lambda version=None: None)
    fr.parser = _mod("freerec.parser", Parser=Parser)
    fr.models = _mod(
        "freerec.models", RecSysArch=RecSysArch, SeqRecArch=SeqRecArch,
        GenRecArch=GenRecArch, PredRecArch=PredRecArch,
    )
    fr.criterions = _mod(
        "freerec.criterions", CrossEntropy4Logits=CrossEntropy4Logits,
        BCELoss4Logits=BCELoss4Logits, BPRLoss=BPRLoss,
        cross_entropy_with_logits=F.cross_entropy,
    )
    fr.launcher = _mod("freerec.launcher", Coach=Coach, EarlyStopError=RuntimeError)
    fr.utils = _mod("freerec.utils", timemeter=timemeter, infoLogger=print, debugLogger=print)
    fr.data = _mod("freerec.data")
    fr.data.fields = _mod("freerec.data.fields", Field=Field)
    fr.data.tags = _mod(
        "freerec.data.tags", USER=USER, ITEM=ITEM, ID=ID, SEQUENCE=SEQUENCE,
        TIMESTAMP=TIMESTAMP, POSITIVE=POSITIVE, NEGATIVE=NEGATIVE, UNSEEN=UNSEEN,
        SEEN=SEEN, LABEL=LABEL, SIZE=SIZE,
    )
    fr.data.datasets = _mod("freerec.data.datasets", RecDataSet=RecDataSet)
    fr.data.datasets.base = _mod("freerec.data.datasets.base", RecDataSet=RecDataSet)
    pp = _mod("freerec.data.postprocessing")
    fr.data.postprocessing = pp
    pp.source = _mod(
        "freerec.data.postprocessing.source",
        RandomShuffledSource=object, OrderedSource=object,
    )
    pp.sampler = _mod("freerec.data.postprocessing.sampler", ValidSampler=object)
    pp.PostProcessor = object

    try:  # torchdata 0.11 dropped datapipes (HSTU/sampler.py:5-11 imports it)
        import torchdata.datapipes  # noqa: F401
    except Exception:
        td = sys.modules.get("torchdata") or _mod("torchdata")
        td.datapipes = _mod(
            "torchdata.datapipes", functional_datapipe=lambda name: (lambda cls: cls)
        )
        td.datapipes.iter = _mod("torchdata.datapipes.iter", IterDataPipe=object)
    return fr


def load_reference(model_dir: str, script: str = "main.py", **cfg_overrides):
    """Import ``/root/reference/<model_dir>/<script>`` unmodified under the shim.

    ``cfg_overrides`` are applied inside ``cfg.compile()`` (the scripts compile their
    parser at import time with an empty argv, SASRec/main.py:28)."""
    install()
    path = REFERENCE_ROOT / model_dir / script
    if not path.exists():
        raise FileNotFoundError(f"{path}: the reference tree is only mounted in the build container")
    for stale in ("modules", "sampler", "quantizer", "converter"):
        sys.modules.pop(stale, None)
    sys.path.insert(0, str(path.parent))
    _PENDING_OVERRIDES.clear()
    _PENDING_OVERRIDES.update(cfg_overrides)
    try:
        if model_dir == "HSTU":  # its sampler.py needs datapipe base classes we do not model
            _mod("sampler", shuffled_time_seqs_source=None)
        spec = importlib.util.spec_from_file_location(f"_ref_{model_dir.replace('-', '_')}", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.pop(0)
        _PENDING_OVERRIDES.clear()
        for stale in ("modules", "sampler"):
            sys.modules.pop(stale, None)
    return mod
