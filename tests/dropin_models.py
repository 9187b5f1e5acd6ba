"""TEST INFRASTRUCTURE -- small self-written models that honour the reference's ``RecSysArch`` contract
(``fit(data) -> {"rec_loss"}``, ``recommend_from_full(data) -> (B,N)``, ``recommend_from_pool``,
``reset_ranking_buffers``, ``encode``; CONTRIBUTING.md:13-16) so that the fused mixins of
``recboard_b200.arch`` can run on the GPU box, where neither /root/reference nor freerec exist.

Every model keeps the reference's hot-path LINES in plain PyTorch (the "eager" methods below are the
statements of SASRec/main.py:195-236, GRU4Rec/main.py:143-178, BERT4Rec/main.py:174-189,
HSTU/main.py:180-209, MF-BPR/main.py:84-109, LightGCN/main.py:77-120); the encoders in front of them are
deliberately tiny stand-ins -- the mixins never look inside ``encode``.
"""
from __future__ import annotations

import types

import torch
import torch.nn as nn
import torch.nn.functional as F

cfg = types.SimpleNamespace(temperature=0.05, num_negs=16)   # HSTU keeps its config in a module-level ``cfg``


class Field(nn.Module):
    """Stand-in for freerec.data.fields.Field: an nn.Module (models attach ``embeddings`` to it), hashable by name."""

    def __init__(self, name: str, count: int = 0):
        super().__init__()
        self.name, self.count = name, count

    def __hash__(self):
        return hash(self.name)

    def __eq__(self, other):
        return isinstance(other, Field) and other.name == self.name


class Arch(nn.Module):
    NUM_PADS = 1
    PADDING_VALUE = 0

    def __init__(self, n_users: int, n_items: int):
        super().__init__()
        self.User, self.Item = Field("User", n_users), Field("Item", n_items)
        self.ISeq, self.IPos, self.INeg = Field("ISeq"), Field("IPos"), Field("INeg")
        self.IUnseen, self.ISeen, self.Size = Field("IUnseen"), Field("ISeen"), Field("Size")
        self.criterion = nn.CrossEntropyLoss(reduction="mean")   # CrossEntropy4Logits(reduction="mean"), SASRec/main.py:126

    @property
    def device(self):
        return next(self.parameters()).device

    def reset_ranking_buffers(self):
        pass

    def forward(self, data, ranking: str = "full"):   # RecSysArch.forward dispatch
        if self.training:
            return self.fit(data)
        if ranking == "full":
            return self.recommend_from_full(data)
        return self.recommend_from_pool(data)


class _SeqEncoder(nn.Module):
    """Embedding + position + one residual MLP block; padding positions are zeroed (the shape of the
    SASRec encoder's output, none of its attention)."""

    def __init__(self, d: int, maxlen: int):
        super().__init__()
        self.pos = nn.Embedding(maxlen, d)
        self.ln = nn.LayerNorm(d)
        self.ff = nn.Sequential(nn.Linear(d, d), nn.GELU(), nn.Linear(d, d))

    def forward(self, x, pad_mask):
        x = x + self.pos.weight[None, -x.shape[1]:, :]
        x = x + self.ff(self.ln(x))
        return x.masked_fill(pad_mask.unsqueeze(-1), 0.0)


class TinySASRec(Arch):
    def __init__(self, n_users, n_items, d=64, maxlen=12):
        super().__init__(n_users, n_items)
        self.Item.add_module("embeddings", nn.Embedding(n_items + self.NUM_PADS, d, padding_idx=self.PADDING_VALUE))
        self.enc = _SeqEncoder(d, maxlen)
        nn.init.normal_(self.Item.embeddings.weight, std=0.3)
        with torch.no_grad():
            self.Item.embeddings.weight[0].zero_()

    def embed(self, seqs):
        return self.Item.embeddings(seqs)                                      # SASRec/main.py:183

    def encode(self, data):
        seqs = data[self.ISeq]
        userEmbds = self.enc(self.embed(seqs), seqs == self.PADDING_VALUE)
        return userEmbds, self.Item.embeddings.weight[self.NUM_PADS:]          # :193

    def fit(self, data):                                                       # SASRec/main.py:195-221 (--loss CE)
        userEmbds, itemEmbds = self.encode(data)
        indices = data[self.ISeq] != self.PADDING_VALUE
        userEmbds = userEmbds[indices]
        logits = torch.einsum("MD,ND->MN", userEmbds, itemEmbds)
        labels = data[self.IPos][indices]
        return {"rec_loss": self.criterion(logits, labels)}

    def recommend_from_full(self, data):                                       # :223-228
        userEmbds, itemEmbds = self.encode(data)
        return torch.einsum("BD,ND->BN", userEmbds[:, -1, :], itemEmbds)

    def recommend_from_pool(self, data):                                       # :230-236
        userEmbds, itemEmbds = self.encode(data)
        return torch.einsum("BD,BKD->BK", userEmbds[:, -1, :], itemEmbds[data[self.IUnseen]])


class TinyGRU4Rec(Arch):
    def __init__(self, n_users, n_items, d=64, maxlen=12):
        super().__init__(n_users, n_items)
        self.Item.add_module("embeddings", nn.Embedding(n_items + self.NUM_PADS, d, padding_idx=self.PADDING_VALUE))
        self.gru = nn.GRU(d, d, batch_first=True)
        nn.init.normal_(self.Item.embeddings.weight, std=0.3)

    def encode(self, data):                                                    # GRU4Rec/main.py:136-152
        seqs = data[self.ISeq]
        out, _ = self.gru(self.Item.embeddings(seqs))
        return out[:, -1, :], self.Item.embeddings.weight[self.NUM_PADS:]      # left-padded: the last step is the last valid one

    def fit(self, data):                                                       # :174-178
        userEmbds, itemEmbds = self.encode(data)
        logits = torch.einsum("BD,ND->BN", userEmbds, itemEmbds)
        return {"rec_loss": self.criterion(logits, data[self.IPos].flatten())}

    def recommend_from_full(self, data):
        userEmbds, itemEmbds = self.encode(data)
        return torch.einsum("BD,ND->BN", userEmbds, itemEmbds)

    def recommend_from_pool(self, data):
        userEmbds, itemEmbds = self.encode(data)
        return torch.einsum("BD,BKD->BK", userEmbds, itemEmbds[data[self.IUnseen]])


class TinyBERT4Rec(Arch):
    NUM_PADS = 2   # pad + [MASK] (BERT4Rec/main.py:60-70)

    def __init__(self, n_users, n_items, d=64, maxlen=12, mask_ratio=0.3):
        super().__init__(n_users, n_items)
        self.mask_ratio, self.MASK_VALUE = mask_ratio, 1
        self.Item.add_module("embeddings", nn.Embedding(n_items + self.NUM_PADS, d, padding_idx=self.PADDING_VALUE))
        self.enc = _SeqEncoder(d, maxlen)
        self.fc = nn.Linear(d, n_items + self.NUM_PADS)
        nn.init.normal_(self.Item.embeddings.weight, std=0.3)
        nn.init.normal_(self.fc.weight, std=0.3)
        nn.init.normal_(self.fc.bias, std=0.2)

    def random_mask(self, seqs, p):                                            # BERT4Rec/main.py:147-160
        padding_mask = seqs == self.PADDING_VALUE
        rnds = torch.rand(seqs.size(), device=seqs.device)
        masks = torch.logical_and(rnds < p, ~padding_mask)
        masked_seqs = seqs.masked_fill(masks, self.MASK_VALUE)
        return masked_seqs, seqs[masks], masks

    def encode(self, data):
        seqs = data[self.ISeq]
        return self.enc(self.Item.embeddings(seqs), seqs == self.PADDING_VALUE)

    def fit(self, data):                                                       # :174-184: (B,S,N+2) GEMM, then the mask
        masked_seqs, labels, masks = self.random_mask(data[self.ISeq], self.mask_ratio)
        data[self.ISeq] = masked_seqs
        logits = self.fc(self.encode(data))[masks]
        return {"rec_loss": self.criterion(logits, labels)}

    def recommend_from_full(self, data):                                       # :186-189
        return self.fc(self.encode(data)[:, -1, :])[:, self.NUM_PADS:]

    def recommend_from_pool(self, data):
        scores = self.fc(self.encode(data)[:, -1, :])[:, self.NUM_PADS:]
        return scores.gather(1, data[self.IUnseen])


class TinyHSTU(TinySASRec):
    """Cosine scores of L2-normalised queries and items; sampled-softmax fit (HSTU/main.py:180-209)."""

    def encode(self, data):
        seqs = data[self.ISeq]
        userEmbds = self.enc(self.embed(seqs), seqs == self.PADDING_VALUE)
        return F.normalize(userEmbds, dim=-1), F.normalize(self.Item.embeddings.weight[self.NUM_PADS:], dim=-1)

    def _sample_negatives(self, userEmbds):                                    # HSTU/main.py:186-190
        g = getattr(self, "_neg_generator", None)
        return torch.randint(0, self.Item.count, (userEmbds.shape[0], cfg.num_negs), device=userEmbds.device, generator=g)

    def fit(self, data):                                                       # :192-202
        userEmbds, itemEmbds = self.encode(data)
        indices = data[self.ISeq] != self.PADDING_VALUE
        userEmbds = userEmbds[indices]
        positives = data[self.IPos][indices].unsqueeze(-1)
        candidates = torch.cat((positives, self._sample_negatives(userEmbds)), dim=1)
        logits = torch.einsum("MD,MKD->MK", userEmbds, itemEmbds[candidates]) / cfg.temperature
        labels = torch.zeros(logits.shape[0], dtype=torch.long, device=logits.device)
        return {"rec_loss": self.criterion(logits, labels)}


class TinyMF(Arch):
    def __init__(self, n_users, n_items, d=64):
        super().__init__(n_users, n_items)
        self.User.add_module("embeddings", nn.Embedding(n_users, d))
        self.Item.add_module("embeddings", nn.Embedding(n_items, d))
        nn.init.normal_(self.User.embeddings.weight, std=0.3)
        nn.init.normal_(self.Item.embeddings.weight, std=0.3)

    def encode(self):
        return self.User.embeddings.weight, self.Item.embeddings.weight

    def fit(self, data):                                                       # MF-BPR/main.py:84-91 (BPR on sampled pairs)
        userEmbds, itemEmbds = self.encode()
        u = userEmbds[data[self.User]]
        pos = torch.einsum("BKD,BKD->BK", u, itemEmbds[data[self.IPos]])
        neg = torch.einsum("BKD,BKD->BK", u, itemEmbds[data[self.INeg]])
        return {"rec_loss": F.softplus(neg - pos).mean()}

    def reset_ranking_buffers(self):                                           # :95-99
        userEmbds, itemEmbds = self.encode()
        self.ranking_buffer = {self.User: userEmbds.detach().clone(), self.Item: itemEmbds.detach().clone()}

    def recommend_from_full(self, data):                                       # :101-104
        userEmbds = self.ranking_buffer[self.User][data[self.User]]
        return torch.einsum("BKD,ND->BN", userEmbds, self.ranking_buffer[self.Item])

    def recommend_from_pool(self, data):                                       # :106-109
        userEmbds = self.ranking_buffer[self.User][data[self.User]]
        return torch.einsum("BKD,BKD->BK", userEmbds, self.ranking_buffer[self.Item][data[self.IUnseen]])


class TinyLightGCN(TinyMF):
    def __init__(self, n_users, n_items, d=64, num_layers=2, nnz_per_user=6, seed=0):
        super().__init__(n_users, n_items, d)
        self.num_layers = num_layers
        g = torch.Generator().manual_seed(seed)
        u = torch.arange(n_users).repeat_interleave(nnz_per_user)
        i = torch.randint(0, n_items, (n_users * nnz_per_user,), generator=g) + n_users
        idx = torch.stack([torch.cat([u, i]), torch.cat([i, u])])
        A = torch.sparse_coo_tensor(idx, torch.ones(idx.shape[1]), (n_users + n_items,) * 2).coalesce()
        deg = torch.sparse.sum(A, dim=1).to_dense().clamp_min(1.0)
        ii = A.indices()
        val = A.values() / torch.sqrt(deg[ii[0]] * deg[ii[1]])                 # symmetric normalisation, LightGCN/main.py:47-49
        self.register_buffer("Adj", torch.sparse_coo_tensor(ii, val, A.shape).coalesce().to_sparse_csr(), persistent=False)

    def encode(self):                                                          # LightGCN/main.py:77-88
        allEmbds = torch.cat((self.User.embeddings.weight, self.Item.embeddings.weight), dim=0)
        avgEmbds = allEmbds / (self.num_layers + 1)
        for _ in range(self.num_layers):
            allEmbds = self.Adj @ allEmbds
            avgEmbds = avgEmbds + allEmbds / (self.num_layers + 1)
        return torch.split(avgEmbds, (self.User.count, self.Item.count))


def lpad_sequences(g, B, S, n_items, pads=1, device="cpu"):
    """(B,S) left-padded id sequences (ids shifted by ``pads``), every row has >= 2 real positions."""
    seq = torch.zeros(B, S, dtype=torch.long)
    for b in range(B):
        n = int(torch.randint(2, S + 1, (1,), generator=g))
        seq[b, S - n:] = torch.randint(0, n_items, (n,), generator=g) + pads
    return seq.to(device)
