"""Drop-in mixins behind the reference's ``RecSysArch`` / ``SeqRecArch`` / ``GenRecArch`` method
contract (``fit(data) -> {"rec_loss": ...}``, ``recommend_from_full(data) -> Tensor[B,N]``,
``reset_ranking_buffers()``; CONTRIBUTING.md:13-16, SASRec/main.py:143-236).

A reference model keeps its ``__init__``, ``encode`` and data pipes; the mixin (listed *before* the
freerec arch class in the bases) replaces only the hot-path lines:

    class SASRecB200(SASRecFused, SASRec): pass          # SASRec from /root/reference/SASRec/main.py

    fit                 lines SASRec/main.py:217-219  -> ops.fused_ce        (no (M,N) logits)
    recommend_from_full lines SASRec/main.py:223-228  -> ops.score_dense     (strict-compat dense path)
    recommend_topk      (new) UniSRec/main.py:408-413 -> ops.topk_eval       (mask + top-K fused)
    recommend_from_pool lines SASRec/main.py:230-236  -> ops.gather_dot      (no (B,K,D) gather)

Every mixin states how the reference picks the query rows (U), the item table view (W), the labels
and the optional bias / temperature; nothing else differs between the six models.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import ops


class FusedFullCatalogMixin:
    """Common machinery.  Subclasses implement ``_train_operands`` and ``_eval_operands``."""

    #: "bf16" | "fp32" | None (None: follow the operand dtype; fp32 tensors -> 3xTF32 "fp32 parity")
    fused_precision: Optional[str] = None
    #: True: the query rows are compacted on the device (``ops.compact_queries``) and their count never travels to the
    #: host -- no ``nonzero()`` synchronisation in ``fit`` (the reference's ``userEmbds[indices]``,
    #: SASRec/main.py:199-200, waits for it every step).  The fused passes are then PLANNED for the capacity B x S and
    #: skip the tiles beyond the count, which under-fills the GPU on small batches; off by default.
    fused_sync_free: bool = False

    # -- hooks -------------------------------------------------------------------------
    def _train_operands(self, data) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, Optional[torch.Tensor], float]:
        """-> (U (M,d), W (N,d), labels (M,), bias (N,)|None, scale)"""
        raise NotImplementedError

    def _eval_operands(self, data) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor], float, int]:
        """-> (U (B,d), W (N',d), bias|None, scale, n_skip) where the first ``n_skip`` rows of W are
        non-item rows (BERT4Rec's pad/mask columns, BERT4Rec/main.py:189)."""
        raise NotImplementedError

    def _whole_table(self, W: torch.Tensor) -> Tuple[torch.Tensor, int]:
        """``W`` is ``self.Item.embeddings.weight[NUM_PADS:]`` itself (SASRec/main.py:193)?  Then hand ``fused_ce`` the
        parameter and the number of pad rows, so that the table gradient comes back as ONE (N+P,d) tensor."""
        weight = getattr(getattr(getattr(self, "Item", None), "embeddings", None), "weight", None)
        P = int(getattr(self, "NUM_PADS", 0))
        if (weight is not None and P > 0 and W.dim() == 2 and W._base is weight and W.is_contiguous()
                and W.shape == (weight.shape[0] - P, weight.shape[1])
                and W.data_ptr() == weight.data_ptr() + P * weight.stride(0) * weight.element_size()):
            return weight, P
        return W, 0

    # -- RecSysArch contract ------------------------------------------------------------
    def fit(self, data: Dict) -> Dict[str, torch.Tensor]:
        U, W, labels, bias, scale, *rest = self._train_operands(data)
        n_valid = rest[0] if rest else None      # device-side row count of a sync-free compaction
        n_skip = 0
        if bias is None:
            W, n_skip = self._whole_table(W)
        return {"rec_loss": ops.fused_ce(U, W, labels, bias=bias, scale=scale, precision=self.fused_precision, n_skip=n_skip,
                                         n_valid=n_valid)}

    def recommend_from_full(self, data: Dict) -> torch.Tensor:
        U, W, bias, scale, n_skip = self._eval_operands(data)
        S = ops.score_dense(U, W, bias=bias, scale=scale, precision=self.fused_precision)
        return S[:, n_skip:] if n_skip else S

    def recommend_from_pool(self, data: Dict) -> torch.Tensor:
        """Scores of the per-row candidate pool ``data[IUnseen]`` (B,K) -- the sampled-ranking protocol
        (SASRec/main.py:230-236, MF-BPR/main.py:106-109) -- through the fused gather-dot: the (B,K,D)
        gathered tensor of the reference is never built."""
        U, W, bias, scale, n_skip = self._eval_operands(data)
        pool = data[self.IUnseen] + n_skip
        S = ops.gather_dot(U, W, pool, scale=scale)
        return S if bias is None else S + bias[pool]

    @torch.no_grad()
    def recommend_topk(self, data: Dict, K: int, seen_crow: Optional[torch.Tensor] = None,
                       seen_col: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Sorted top-K (vals, item ids) per row with the row's seen items skipped -- the fused
        stand-in for ``scores = model(data, ranking="full"); scores[seen] = -1e23; topk``
        (UniSRec/main.py:408-413).  ``seen`` is the CSR of ``data[ISeen]`` over 0-based item ids."""
        U, W, bias, scale, n_skip = self._eval_operands(data)
        if n_skip:
            W = W[n_skip:]
            bias = bias[n_skip:].contiguous() if bias is not None else None
        return ops.topk_eval(U, W, K, seen_crow, seen_col, bias=bias, scale=scale, precision=self.fused_precision)


class SASRecFused(FusedFullCatalogMixin):
    """SASRec ``--loss CE``: queries = every non-pad position, labels = IPos there
    (SASRec/main.py:197-200,217-219); eval query = last position (:226-228)."""

    def _train_operands(self, data):
        userEmbds, itemEmbds = self.encode(data)
        indices = data[self.ISeq] != self.PADDING_VALUE
        if self.fused_sync_free:
            U, (labels,), count = ops.compact_queries(userEmbds, indices, data[self.IPos])
            return U, itemEmbds, labels, None, 1.0, count
        return userEmbds[indices], itemEmbds, data[self.IPos][indices], None, 1.0

    def _eval_operands(self, data):
        userEmbds, itemEmbds = self.encode(data)
        return userEmbds[:, -1, :], itemEmbds, None, 1.0, 0


class GRU4RecFused(FusedFullCatalogMixin):
    """GRU4Rec ``--loss CE``: one query per sequence (last valid step, GRU4Rec/main.py:143-150),
    labels = IPos.flatten() (:174-178)."""

    def _train_operands(self, data):
        userEmbds, itemEmbds = self.encode(data)
        return userEmbds, itemEmbds, data[self.IPos].flatten(), None, 1.0

    def _eval_operands(self, data):
        userEmbds, itemEmbds = self.encode(data)
        return userEmbds, itemEmbds, None, 1.0, 0


class BERT4RecFused(FusedFullCatalogMixin):
    """BERT4Rec: ``nn.Linear(D, N+2)`` head with bias, CE over all N+2 columns at the masked
    positions (BERT4Rec/main.py:174-184).  The reference runs the (B,S,N+2) GEMM *before* masking
    (:181); here the masked rows are selected first and only they are scored."""

    def _train_operands(self, data):
        seqs = data[self.ISeq]
        masked_seqs, labels, masks = self.random_mask(seqs=seqs, p=self.mask_ratio)
        data[self.ISeq] = masked_seqs
        userEmbds = self.encode(data)
        if self.fused_sync_free:   # the labels are the original ids at the masked positions (BERT4Rec/main.py:147-160);
            # (the model's own random_mask still indexes ``seqs[masks]``: that synchronisation is inside the reference method)
            U, (lab,), count = ops.compact_queries(userEmbds, masks, seqs)
            return U, self.fc.weight, lab, self.fc.bias, 1.0, count
        return userEmbds[masks], self.fc.weight, labels, self.fc.bias, 1.0

    def _eval_operands(self, data):
        userEmbds = self.encode(data)
        return userEmbds[:, -1, :], self.fc.weight, self.fc.bias, 1.0, self.NUM_PADS


class HSTUFused(FusedFullCatalogMixin):
    """HSTU full ranking: cosine scores of L2-normalised queries and items (HSTU/main.py:180-184,
    204-209).  The reference re-normalises the whole (N,d) table inside ``encode`` for every batch;
    here ``reset_ranking_buffers`` builds the normalised operand copy once per evaluation sweep
    (``ops.normalize_rows``, one HBM pass, optionally bf16) and ``_eval_operands`` reuses it.  A model
    that wants to skip the redundant per-batch table pass as well overrides ``encode_users``.
    ``fit`` is sampled softmax over 1 positive + ``num_negs`` sampled negatives (HSTU/main.py:186-202): the
    reference gathers an (M, 1+num_negs, D) tensor (~200 MB per step at the Beauty shape) and contracts it
    row-wise; here ``ops.gather_dot`` scores the ids in place."""

    #: softmax temperature; None = ``cfg.temperature`` of the reference module the model class comes from
    fused_temperature: Optional[float] = None

    def _temperature(self) -> float:
        if self.fused_temperature is not None:
            return float(self.fused_temperature)
        for klass in type(self).__mro__:   # the reference keeps its config in a module-level ``cfg`` (HSTU/main.py:16-40)
            fn = vars(klass).get("_sample_negatives") or vars(klass).get("encode")
            cfg = getattr(fn, "__globals__", {}).get("cfg")
            if cfg is not None and hasattr(cfg, "temperature"):
                return float(cfg.temperature)
        raise AttributeError("set fused_temperature: no cfg.temperature found for this model")

    def fit(self, data):
        userEmbds, itemEmbds = self.encode(data)
        indices = data[self.ISeq] != self.PADDING_VALUE
        userEmbds = userEmbds[indices]                                              # (M, D)
        positives = data[self.IPos][indices].unsqueeze(-1)                          # (M, 1)
        candidates = torch.cat((positives, self._sample_negatives(userEmbds)), dim=1)   # (M, 1 + num_negs)
        logits = ops.gather_dot(userEmbds, itemEmbds, candidates, scale=1.0 / self._temperature())
        labels = torch.zeros(logits.shape[0], dtype=torch.long, device=logits.device)   # the positive sits in column 0
        return {"rec_loss": self.criterion(logits, labels)}

    def reset_ranking_buffers(self):
        sup = super()
        if hasattr(sup, "reset_ranking_buffers"):
            sup.reset_ranking_buffers()
        out_dtype = torch.bfloat16 if self.fused_precision == "bf16" else None
        w = self.Item.embeddings.weight
        self._fused_item = ops.normalize_rows(w[self.NUM_PADS:], out_dtype=out_dtype)
        self._fused_item_key = (w.data_ptr(), w._version)   # the cache is only valid for these exact weights

    def encode_users(self, data) -> torch.Tensor:
        """(B, S, d) normalised user states; default = the reference's ``encode`` (HSTU/main.py:164-184)."""
        return self.encode(data)[0]

    def _eval_operands(self, data):
        w = self.Item.embeddings.weight
        stale = getattr(self, "_fused_item_key", None) != (w.data_ptr(), w._version)   # optimizer step / load_state_dict since
        if getattr(self, "_fused_item", None) is None or self.training or stale:
            userEmbds, itemEmbds = self.encode(data)   # the reference's own per-call normalisation (HSTU/main.py:180-184)
            return userEmbds[:, -1, :], itemEmbds, None, 1.0, 0
        return self.encode_users(data)[:, -1, :], self._fused_item, None, 1.0, 0


class GenRecFused(FusedFullCatalogMixin):
    """MF-BPR / LightGCN family: eval queries are rows of ``ranking_buffer[User]`` picked by
    ``data[User]`` (B,1) (MF-BPR/main.py:95-104, LightGCN/main.py:110-120).  ``fit`` (BPR on sampled
    pairs) is not a full-catalog contraction and stays with the reference."""

    def fit(self, data):
        return super(FusedFullCatalogMixin, self).fit(data)

    def reset_ranking_buffers(self):
        super().reset_ranking_buffers()
        # build the operand copies the fused eval wants once per sweep, not once per batch
        self._fused_user = self.ranking_buffer[self.User].contiguous()
        self._fused_item = self.ranking_buffer[self.Item].contiguous()
        if self.fused_precision == "bf16":
            self._fused_item = self._fused_item.to(torch.bfloat16)

    def _eval_operands(self, data):
        users = data[self.User]
        U = ops.gather_rows_raw(self._fused_user, users.reshape(-1).contiguous())  # "BKD" with K == 1
        return U, self._fused_item, None, 1.0, 0


class LightGCNFused(GenRecFused):
    """LightGCN: the propagation in front of the ranking buffers, ``allEmbds = self.Adj @ allEmbds`` x L with
    the layer average (LightGCN/main.py:77-88), runs through ``ops.spmm`` (rb_spmm_csr; the symmetric
    normalisation makes the backward the same product).  Everything after it is ``GenRecFused``."""

    def encode(self):
        allEmbds = torch.cat((self.User.embeddings.weight, self.Item.embeddings.weight), dim=0)
        avgEmbds = allEmbds / (self.num_layers + 1)
        for _ in range(self.num_layers):
            allEmbds = ops.spmm(self.Adj, allEmbds, symmetric=True)
            avgEmbds = avgEmbds + allEmbds / (self.num_layers + 1)
        return torch.split(avgEmbds, (self.User.count, self.Item.count))


class FusedEmbedding(torch.nn.Embedding):
    """``nn.Embedding`` whose lookup and dense backward run through ``rb_gather_rows`` / the one-launch deterministic
    scatter-add (``self.Item.embeddings(seqs)``, SASRec/main.py:183, and its autograd at :249).  With
    ``accumulate_grad=True`` the backward adds the rows straight into ``weight.grad`` once that buffer exists
    (``zero_grad(set_to_none=False)``): no fresh dense (N+P,d) gradient per step."""

    accumulate_grad: bool = False

    def forward(self, idx: torch.Tensor) -> torch.Tensor:
        pad = -1 if self.padding_idx is None else int(self.padding_idx)
        return ops.gather_rows(self.weight, idx, padding_idx=pad, accumulate=self.accumulate_grad)


def fuse_item_embedding(model, accumulate_grad: bool = False):
    """Swap ``model.Item.embeddings`` for a ``FusedEmbedding`` that shares the same weight Parameter (state_dict keys
    and optimizer state are unaffected)."""
    old = model.Item.embeddings
    new = FusedEmbedding(old.num_embeddings, old.embedding_dim, padding_idx=old.padding_idx, device="meta")
    new.weight = old.weight
    new.accumulate_grad = accumulate_grad
    model.Item.embeddings = new
    return model


def normalized_table(weight: torch.Tensor, num_pads: int = 1, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """``F.normalize(weight[NUM_PADS:], dim=-1)`` (HSTU/main.py:182-184) through ``rb_normalize_rows`` --
    for callers that cache the normalised table across an evaluation sweep."""
    return ops.normalize_rows(weight[num_pads:], out_dtype=out_dtype)
