"""TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/*.npz`` by running the reference's
own, unmodified model files (imported from /root/reference through ``oracle/freerec_shim``)
on small seeded inputs, and recording the tensors that cross the hot-path boundary:

    U (query rows), W (item table view), bias, labels  ->  scores / loss / dU / dW / dbias
    gather indices + upstream grad                     ->  embedding-table gradient

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py

The reference code is *instrumented from outside* (forward hooks / wrapper around
``encode``), never edited.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import freerec_shim as shim  # noqa: E402

OUT = ROOT / "tests" / "golden"


def _np(t):
    return t.detach().cpu().numpy()


def _seqs(g, B, S, N, pads=1, min_len=1):
    """Left-padded item sequences with ids offset by NUM_PADS (lpad_, SASRec/main.py:150-154)."""
    lens = torch.randint(min_len, S + 1, (B,), generator=g)
    seq = torch.zeros(B, S, dtype=torch.long)
    for b in range(B):
        L = int(lens[b])
        seq[b, S - L:] = torch.randint(0, N, (L,), generator=g) + pads
    return seq


def _capture_encode(model):
    """Wrap model.encode so its outputs are cached with retain_grad (no source edits)."""
    cache = {}
    orig = model.encode

    def wrapped(*a, **k):
        out = orig(*a, **k)
        outs = out if isinstance(out, tuple) else (out,)
        for o in outs:
            if o.requires_grad and not o.is_leaf:
                o.retain_grad()
        cache["out"] = outs
        return out

    model.encode = wrapped
    return cache


def gen_sasrec():
    N, d, B, S = 300, 64, 16, 12
    torch.manual_seed(2026)
    ref = shim.load_reference("SASRec", loss="CE", embedding_dim=d, maxlen=S, dropout_rate=0.0)
    ds = shim.RecDataSet(n_users=B, n_items=N)
    model = ref.SASRec(ds)
    g = torch.Generator().manual_seed(11)
    ISeq = _seqs(g, B, S, N)
    IPos = torch.randint(0, N, (B, S), generator=g)
    IPos[ISeq == 0] = 0
    data = {model.ISeq: ISeq, model.IPos: IPos}

    cache = _capture_encode(model)
    model.train()
    loss = model(data)["rec_loss"]
    userEmbds, itemEmbds = cache["out"]
    loss.backward()
    indices = ISeq != 0
    U = userEmbds[indices]
    dU = userEmbds.grad[indices]
    dTable = model.Item.embeddings.weight.grad
    labels = IPos[indices]

    model.eval()
    with torch.no_grad():
        scores = model(data, ranking="full")
        uE, iE = model.encode(data)
    np.savez_compressed(
        OUT / "sasrec_ce.npz",
        ISeq=_np(ISeq), IPos=_np(IPos), U=_np(U), W=_np(itemEmbds), labels=_np(labels),
        loss=_np(loss), dU=_np(dU), dTable_total=_np(dTable),
        table=_np(model.Item.embeddings.weight),
        U_eval=_np(uE[:, -1, :]), scores_full=_np(scores),
    )
    print("sasrec_ce: loss", loss.item(), "M", U.shape[0])


def gen_gru4rec():
    N, d, B, S = 257, 64, 24, 10
    torch.manual_seed(2027)
    ref = shim.load_reference("GRU4Rec", loss="CE", embedding_dim=d, maxlen=S)
    ds = shim.RecDataSet(n_users=B, n_items=N)
    model = ref.GRU4Rec(ds)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    g = torch.Generator().manual_seed(12)
    ISeq = _seqs(g, B, S, N)
    IPos = torch.randint(0, N, (B, 1), generator=g)
    INeg = torch.randint(0, N, (B, 1), generator=g)
    data = {model.ISeq: ISeq, model.IPos: IPos, model.INeg: INeg}
    cache = _capture_encode(model)
    model.train()
    loss = model(data)["rec_loss"]
    loss.backward()
    userEmbds, itemEmbds = cache["out"]
    model.eval()
    with torch.no_grad():
        scores = model(data, ranking="full")
        uE, _ = model.encode(data)
    np.savez_compressed(
        OUT / "gru4rec_ce.npz",
        U=_np(userEmbds), W=_np(itemEmbds), labels=_np(IPos.flatten()), loss=_np(loss),
        dU=_np(userEmbds.grad), U_eval=_np(uE), scores_full=_np(scores),
    )
    print("gru4rec_ce: loss", loss.item())


def gen_bert4rec():
    N, d, B, S = 203, 64, 12, 16
    torch.manual_seed(2028)
    ref = shim.load_reference(
        "BERT4Rec", embedding_dim=d, maxlen=S, dropout_rate=0.0, num_heads=2, num_blocks=1
    )
    ds = shim.RecDataSet(n_users=B, n_items=N)
    model = ref.BERT4Rec(ds)
    with torch.no_grad():  # non-zero bias so the bias path is really exercised
        model.fc.bias.normal_(0.0, 0.1)
        model.fc.weight.mul_(20.0)
    g = torch.Generator().manual_seed(13)
    ISeq = _seqs(g, B, S, N, pads=2, min_len=4)
    data = {model.ISeq: ISeq.clone()}
    cache = _capture_encode(model)
    captured = {}
    orig_mask = model.random_mask

    def mask_wrapped(*a, **k):
        out = orig_mask(*a, **k)
        captured["masked_seqs"], captured["labels"], captured["masks"] = out
        return out

    model.random_mask = mask_wrapped
    model.train()
    loss = model(data)["rec_loss"]
    loss.backward()
    (userEmbds,) = cache["out"]
    masks = captured["masks"]
    np.savez_compressed(
        OUT / "bert4rec_ce.npz",
        U=_np(userEmbds[masks]), W=_np(model.fc.weight), bias=_np(model.fc.bias),
        labels=_np(captured["labels"]), loss=_np(loss), dU=_np(userEmbds.grad[masks]),
        dW=_np(model.fc.weight.grad), dbias=_np(model.fc.bias.grad),
    )
    model.eval()
    with torch.no_grad():
        data = {model.ISeq: ISeq.clone()}
        scores = model(data, ranking="full")
        uE = model.encode(data)[:, -1, :]
    np.savez_compressed(
        OUT / "bert4rec_full.npz",
        U=_np(uE), W=_np(model.fc.weight), bias=_np(model.fc.bias), scores_full=_np(scores),
        num_pads=np.int64(2),
    )
    print("bert4rec_ce: loss", loss.item(), "M", int(masks.sum()))


def gen_mf_lightgcn():
    U_, N, d, B = 96, 411, 64, 32
    for name, cls in (("MF-BPR", "MF"), ("LightGCN", "LightGCN")):
        torch.manual_seed(2029)
        ref = shim.load_reference(name, embedding_dim=d)
        ds = shim.RecDataSet(n_users=U_, n_items=N)
        model = getattr(ref, cls)(ds)
        with torch.no_grad():
            model.User.embeddings.weight.normal_(0, 0.3)
            model.Item.embeddings.weight.normal_(0, 0.3)
        model.eval()
        g = torch.Generator().manual_seed(14)
        users = torch.randperm(U_, generator=g)[:B].unsqueeze(1)
        with torch.no_grad():
            model.reset_ranking_buffers()
            scores = model({model.User: users}, ranking="full")
        np.savez_compressed(
            OUT / f"{cls.lower()}_full.npz",
            users=_np(users), user_table=_np(model.ranking_buffer[model.User]),
            item_table=_np(model.ranking_buffer[model.Item]), scores_full=_np(scores),
        )
        print(f"{cls}_full: scores", tuple(scores.shape))


def gen_hstu():
    N, d, B, S = 222, 64, 8, 12
    torch.manual_seed(2030)
    ref = shim.load_reference("HSTU", embedding_dim=d, maxlen=S)
    ds = shim.RecDataSet(n_users=B, n_items=N)
    model = ref.HSTU(ds)
    model.eval()
    g = torch.Generator().manual_seed(15)
    ISeq = _seqs(g, B, S, N, min_len=S)
    Time = torch.sort(torch.randint(0, 10_000_000, (B, S), generator=g), dim=1).values
    data = {model.ISeq: ISeq, model.Time: Time}
    with torch.no_grad():
        scores = model(data, ranking="full")
        uE, iE = model.encode(data)
    np.savez_compressed(
        OUT / "hstu_full.npz", U=_np(uE[:, -1, :]), W=_np(iE), scores_full=_np(scores),
        table=_np(model.Item.embeddings.weight),
    )
    print("hstu_full: scores", tuple(scores.shape))


def gen_gather_backward():
    """nn.Embedding(padding_idx=0) forward/backward exactly as the reference instantiates it
    (SASRec/main.py:70-77,183) with duplicate and pad ids."""
    torch.manual_seed(2031)
    N, d = 97, 32
    emb = torch.nn.Embedding(N + 1, d, padding_idx=0)
    g = torch.Generator().manual_seed(16)
    idx = torch.randint(0, 12, (9, 7), generator=g)  # few distinct ids => many duplicates
    idx[0, :3] = 0
    out = emb(idx)
    go = torch.randn(out.shape, generator=g)
    out.backward(go)
    np.savez_compressed(
        OUT / "embedding_bwd.npz", table=_np(emb.weight), idx=_np(idx), out=_np(out),
        grad_out=_np(go), grad_table=_np(emb.weight.grad),
    )
    print("embedding_bwd ok")


def gen_pool_and_sampled():
    """pool ranking of SASRec (SASRec/main.py:230-236) and HSTU (cosine), HSTU's table
    normalisation (HSTU/main.py:180-184) and its sampled-softmax fit (HSTU/main.py:186-202) with the
    negatives recorded, so the fused gather-dot can be checked against the reference's own numbers."""
    N, d, B, S = 150, 32, 8, 10
    torch.manual_seed(2032)
    ref = shim.load_reference("SASRec", loss="CE", embedding_dim=d, maxlen=S, dropout_rate=0.0)
    model = ref.SASRec(shim.RecDataSet(n_users=B, n_items=N))
    model.eval()
    g = torch.Generator().manual_seed(17)
    ISeq = _seqs(g, B, S, N)
    pool = torch.randint(0, N, (B, 21), generator=g)
    data = {model.ISeq: ISeq, model.IUnseen: pool}
    with torch.no_grad():
        uE, iE = model.encode(data)
        scores = model(data, ranking="pool")
    np.savez_compressed(OUT / "sasrec_pool.npz", U=_np(uE[:, -1, :]), W=_np(iE), pool=_np(pool), scores_pool=_np(scores))
    print("sasrec_pool: scores", tuple(scores.shape))

    torch.manual_seed(2033)
    ref = shim.load_reference("HSTU", embedding_dim=d, maxlen=S)
    model = ref.HSTU(shim.RecDataSet(n_users=B, n_items=N))
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    g = torch.Generator().manual_seed(18)
    ISeq = _seqs(g, B, S, N)
    Time = torch.sort(torch.randint(0, 10_000_000, (B, S), generator=g), dim=1).values
    IPos = torch.randint(0, N, (B, S), generator=g)
    model.eval()
    with torch.no_grad():
        data = {model.ISeq: ISeq, model.Time: Time, model.IUnseen: pool}
        uE, iE = model.encode(data)
        pool_scores = model(data, ranking="pool")
    # the sampled-softmax fit, with the negatives it drew recorded from outside
    model.train()
    drawn = {}
    orig = model._sample_negatives

    def recording(userEmbds):
        drawn["neg"] = orig(userEmbds)
        return drawn["neg"]

    model._sample_negatives = recording
    cache = _capture_encode(model)
    data = {model.ISeq: ISeq, model.Time: Time, model.IPos: IPos}
    loss = model(data)["rec_loss"]
    loss.backward()
    uE_t, iE_t = cache["out"]
    indices = ISeq != 0
    np.savez_compressed(
        OUT / "hstu_sampled.npz", table=_np(model.Item.embeddings.weight), W_norm=_np(iE),
        U_pool=_np(uE[:, -1, :]), pool=_np(pool), scores_pool=_np(pool_scores),
        U_fit=_np(uE_t[indices]), W_fit=_np(iE_t), positives=_np(IPos[indices]), negatives=_np(drawn["neg"]),
        temperature=np.float32(ref.cfg.temperature), loss=_np(loss),
        dU_fit=_np(uE_t.grad[indices]), dW_fit=_np(iE_t.grad),
    )
    print("hstu_sampled: loss", float(loss), "negatives", tuple(drawn["neg"].shape))


def gen_lightgcn_propagation():
    """LightGCN.encode (LightGCN/main.py:77-88) -- adjacency in CSR, tables in, propagated tables out."""
    U_, N, d = 60, 140, 32
    torch.manual_seed(2034)
    ref = shim.load_reference("LightGCN", embedding_dim=d)
    model = ref.LightGCN(shim.RecDataSet(n_users=U_, n_items=N))
    with torch.no_grad():
        model.User.embeddings.weight.normal_(0, 0.3)
        model.Item.embeddings.weight.normal_(0, 0.3)
        uE, iE = model.encode()
    A = model.Adj
    np.savez_compressed(
        OUT / "lightgcn_prop.npz", crow=_np(A.crow_indices()), col=_np(A.col_indices()), val=_np(A.values()),
        user_table=_np(model.User.embeddings.weight), item_table=_np(model.Item.embeddings.weight),
        num_layers=np.int64(model.num_layers), user_out=_np(uE), item_out=_np(iE),
    )
    print("lightgcn_prop: nnz", int(A.values().numel()), "layers", int(model.num_layers))


def gen_unisrec_evaluate():
    """the in-tree statement of ``Coach.evaluate`` (UniSRec/main.py:400-447), run UNMODIFIED with a
    stub ``self`` whose ``monitor`` records what the reference hands to the metric functions: the seen-masked dense
    ``scores`` (``scores[seen] = -1e23``, :409-413) and the dense multi-hot ``targets`` (:414), per batch, together
    with the batch size it weights them by (``n=bsz``, :403,430).  Two batches (the second ragged), rows with empty
    seen lists, a target inside the seen list, multi-target rows."""
    import types
    ref = shim.load_reference("UniSRec")
    N, d = 211, 16
    g = torch.Generator().manual_seed(2035)
    W = torch.randn(N, d, generator=g)
    Item = shim.Field("Item", N)
    ISeen, IUnseen, Size = shim.Field("ISeen"), shim.Field("IUnseen"), shim.Field("Size")
    batches, raw = [], []
    for b, B in enumerate((24, 13)):
        U = torch.randn(B, d, generator=g)
        seen = [torch.randperm(N, generator=g)[: int(torch.randint(0, 12, (1,), generator=g))].tolist() for _ in range(B)]
        seen[0] = []
        unseen = [torch.randperm(N, generator=g)[: (3 if r % 5 == 4 else 1)].tolist() for r in range(B)]
        if seen[1]:
            unseen[1] = [seen[1][0]]          # a target the user has already seen: masked before ranking, never hit
        scores = U @ W.T
        batches.append({"dataset": "D", "scores": scores, ISeen: seen, IUnseen: unseen, Size: B})
        raw.append((U, seen, unseen))
    recorded = []

    class _Self:
        pass

    me = _Self()
    me.dataloader = batches
    me.Size, me.ISeen, me.IUnseen = Size, ISeen, IUnseen
    me.device = torch.device("cpu")
    me.remove_seen = True
    me.cfg = types.SimpleNamespace(ranking="full")
    me.datasets = {"D": types.SimpleNamespace(fields=shim._FieldTable({shim.ITEM: Item}))}
    me.get_res_sys_arch = lambda: types.SimpleNamespace(reset_ranking_buffers=lambda: None)
    me.dict_to_device = lambda data: data
    me.model = lambda data, ranking="full": data["scores"].clone()
    me.monitor = lambda scores, targets, n, reduction, mode, pool: recorded.append((scores.clone(), targets.clone(), n, list(pool)))
    ref.CoachForUniSRec.evaluate(me, epoch=0, mode="test")
    out = {"item_table": _np(W), "n_batches": np.int64(len(batches))}
    for b, (U, seen, unseen) in enumerate(raw):
        s_m, t_d, n, pool = recorded[2 * b]                     # two monitor calls per batch (:428-447), same tensors
        assert torch.equal(recorded[2 * b + 1][0], s_m) and pool == ["HITRATE", "PRECISION", "RECALL", "NDCG", "MRR"]
        crow, col = [0], []
        for r in seen:
            col += sorted(set(r)); crow.append(len(col))
        tcrow, tcol = [0], []
        for r in unseen:
            tcol += sorted(set(r)); tcrow.append(len(tcol))
        out.update({f"U{b}": _np(U), f"seen_crow{b}": np.asarray(crow, np.int64), f"seen_col{b}": np.asarray(col, np.int64),
                    f"tgt_crow{b}": np.asarray(tcrow, np.int64), f"tgt_col{b}": np.asarray(tcol, np.int64),
                    f"scores_masked{b}": _np(s_m), f"targets{b}": _np(t_d), f"bsz{b}": np.int64(n)})
    np.savez_compressed(OUT / "unisrec_evaluate.npz", **out)
    print("unisrec_evaluate: batches", len(batches), "masked entries", int(sum((r[0] == -1e23).sum() for r in recorded[::2])))


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(1)  # deterministic summation order for the committed vectors
    gen_sasrec()
    gen_gru4rec()
    gen_bert4rec()
    gen_mf_lightgcn()
    gen_hstu()
    gen_gather_backward()
    gen_pool_and_sampled()
    gen_lightgcn_propagation()
    gen_unisrec_evaluate()


if __name__ == "__main__":
    main()
