"""Drop-in boundary, host side: the six reference models, imported UNMODIFIED from /root/reference
through the freerec shim, get the fused mixins; with the CUDA ops replaced by the oracle (this test
runs without a GPU) the mixin's choice of query rows / table view / labels / bias must reproduce the
reference's own ``fit`` and ``recommend_from_full`` outputs.  Skipped where /root/reference is absent
(the GPU box); the kernels themselves are covered by tests/test_gpu_parity.py."""
import types
from pathlib import Path

import pytest
import torch

from oracle import freerec_shim as shim
from oracle import reference_path as orc
from recboard_b200 import arch

pytestmark = pytest.mark.skipif(not Path("/root/reference/SASRec/main.py").exists(), reason="reference tree not mounted")


@pytest.fixture()
def oracle_ops(monkeypatch):
    """Stand-in for recboard_b200.ops on a GPU-less host (test only)."""
    stub = types.SimpleNamespace(
        fused_ce=lambda U, W, labels, bias=None, scale=1.0, precision=None, n_skip=0, n_valid=None: orc.ce_loss(U, W[n_skip:], labels, bias, scale),
        score_dense=lambda U, W, bias=None, scale=1.0, precision=None: orc.score_dense(U, W, bias, scale),
        gather_rows_raw=lambda table, idx: table[idx],
        spmm=lambda A, X, symmetric=False, At=None: orc.spmm(A, X),
        gather_dot=lambda U, table, idx, scale=1.0, padding_idx=-1: orc.gather_dot(U, table, idx, scale),
        normalize_rows=lambda x, out_dtype=None, eps=1e-12: orc.normalize_rows(x, eps),
        topk_eval=lambda U, W, K, crow=None, col=None, bias=None, scale=1.0, precision=None:
            orc.topk_sorted(orc.mask_seen(orc.score_dense(U, W, bias, scale), crow, col) if crow is not None
                            else orc.score_dense(U, W, bias, scale), K),
    )
    monkeypatch.setattr(arch, "ops", stub)
    return stub


def _seqs(g, B, S, N, pads=1):
    seq = torch.zeros(B, S, dtype=torch.long)
    for b in range(B):
        L = int(torch.randint(2, S + 1, (1,), generator=g))
        seq[b, S - L:] = torch.randint(0, N, (L,), generator=g) + pads
    return seq


def test_sasrec_fit_and_full(oracle_ops):
    ref = shim.load_reference("SASRec", loss="CE", embedding_dim=32, maxlen=10, dropout_rate=0.0)
    Fused = type("SASRecB200", (arch.SASRecFused, ref.SASRec), {})
    torch.manual_seed(0)
    model = Fused(shim.RecDataSet(n_users=8, n_items=120))
    g = torch.Generator().manual_seed(1)
    ISeq = _seqs(g, 8, 10, 120)
    IPos = torch.randint(0, 120, (8, 10), generator=g)
    data = {model.ISeq: ISeq, model.IPos: IPos}
    model.train()
    fused = model(data)["rec_loss"]
    plain = ref.SASRec.fit(model, data)["rec_loss"]
    assert torch.allclose(fused, plain, rtol=1e-6)
    model.eval()
    with torch.no_grad():
        assert torch.allclose(model(data, ranking="full"), ref.SASRec.recommend_from_full(model, data), rtol=1e-6, atol=1e-6)
        vals, ids = model.recommend_topk(data, 5)
        rv, ri = orc.topk_sorted(ref.SASRec.recommend_from_full(model, data), 5)
        assert torch.equal(ids, ri)
        pool = {**data, model.IUnseen: torch.randint(0, 120, (8, 11), generator=g)}
        assert torch.allclose(model(pool, ranking="pool"), ref.SASRec.recommend_from_pool(model, pool), rtol=1e-6, atol=1e-6)


def test_gru4rec_fit(oracle_ops):
    ref = shim.load_reference("GRU4Rec", loss="CE", embedding_dim=32, maxlen=10)
    Fused = type("GRU4RecB200", (arch.GRU4RecFused, ref.GRU4Rec), {})
    torch.manual_seed(0)
    model = Fused(shim.RecDataSet(n_users=8, n_items=90))
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    g = torch.Generator().manual_seed(2)
    data = {model.ISeq: _seqs(g, 8, 10, 90), model.IPos: torch.randint(0, 90, (8, 1), generator=g),
            model.INeg: torch.randint(0, 90, (8, 1), generator=g)}
    model.train()
    assert torch.allclose(model(data)["rec_loss"], ref.GRU4Rec.fit(model, data)["rec_loss"], rtol=1e-6)


def test_bert4rec_fit_and_full(oracle_ops):
    ref = shim.load_reference("BERT4Rec", embedding_dim=32, maxlen=12, dropout_rate=0.0, num_heads=2, num_blocks=1)
    Fused = type("BERT4RecB200", (arch.BERT4RecFused, ref.BERT4Rec), {})
    torch.manual_seed(0)
    model = Fused(shim.RecDataSet(n_users=6, n_items=70))
    with torch.no_grad():
        model.fc.bias.normal_(0, 0.2)
    g = torch.Generator().manual_seed(3)
    ISeq = _seqs(g, 6, 12, 70, pads=2)
    model.train()
    torch.manual_seed(42)
    fused = model({model.ISeq: ISeq.clone()})["rec_loss"]
    torch.manual_seed(42)  # same random mask
    plain = ref.BERT4Rec.fit(model, {model.ISeq: ISeq.clone()})["rec_loss"]
    assert torch.allclose(fused, plain, rtol=1e-5)
    model.eval()
    with torch.no_grad():
        a = model({model.ISeq: ISeq.clone()}, ranking="full")
        b = ref.BERT4Rec.recommend_from_full(model, {model.ISeq: ISeq.clone()})
        assert a.shape == b.shape == (6, 70) and torch.allclose(a, b, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name,cls", [("MF-BPR", "MF"), ("LightGCN", "LightGCN")])
def test_genrec_full(oracle_ops, name, cls):
    ref = shim.load_reference(name, embedding_dim=32)
    Base = getattr(ref, cls)
    Fused = type(cls + "B200", (arch.GenRecFused, Base), {})
    torch.manual_seed(0)
    model = Fused(shim.RecDataSet(n_users=40, n_items=150))
    with torch.no_grad():
        model.User.embeddings.weight.normal_(0, 0.3)
        model.Item.embeddings.weight.normal_(0, 0.3)
    model.eval()
    users = torch.arange(0, 40, 3).unsqueeze(1)
    with torch.no_grad():
        model.reset_ranking_buffers()
        a = model({model.User: users}, ranking="full")
        b = Base.recommend_from_full(model, {model.User: users})
    assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)
    # BPR fit is left to the reference
    assert Fused.fit is not Base.fit and "super(FusedFullCatalogMixin" in __import__("inspect").getsource(Fused.fit)


def test_hstu_full(oracle_ops):
    ref = shim.load_reference("HSTU", embedding_dim=32, maxlen=10)
    Fused = type("HSTUB200", (arch.HSTUFused, ref.HSTU), {})
    torch.manual_seed(0)
    model = Fused(shim.RecDataSet(n_users=4, n_items=80))
    model.eval()
    g = torch.Generator().manual_seed(4)
    data = {model.ISeq: _seqs(g, 4, 10, 80), model.Time: torch.sort(torch.randint(0, 10**6, (4, 10), generator=g), 1).values}
    with torch.no_grad():
        assert torch.allclose(model(data, ranking="full"), ref.HSTU.recommend_from_full(model, data), rtol=1e-6, atol=1e-7)
        pool = {**data, model.IUnseen: torch.randint(0, 80, (4, 9), generator=g)}
        assert torch.allclose(model(pool, ranking="pool"), ref.HSTU.recommend_from_pool(model, pool), rtol=1e-6, atol=1e-7)
        model.reset_ranking_buffers()  # normalised table cached once per sweep (a11)
        assert torch.allclose(model(data, ranking="full"), ref.HSTU.recommend_from_full(model, data), rtol=1e-6, atol=1e-7)
        assert model._fused_item.shape == (80, 32)


def test_hstu_sampled_softmax_fit(oracle_ops):
    """the fused gather-dot fit equals the reference's gather + einsum fit (same sampled negatives)."""
    ref = shim.load_reference("HSTU", embedding_dim=32, maxlen=10)
    Fused = type("HSTUB200", (arch.HSTUFused, ref.HSTU), {})
    torch.manual_seed(0)
    model = Fused(shim.RecDataSet(n_users=4, n_items=80))
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.train()
    g = torch.Generator().manual_seed(5)
    data = {model.ISeq: _seqs(g, 4, 10, 80), model.IPos: torch.randint(0, 80, (4, 10), generator=g),
            model.Time: torch.sort(torch.randint(0, 10**6, (4, 10), generator=g), 1).values}
    torch.manual_seed(7)
    fused = model(data)["rec_loss"]
    torch.manual_seed(7)   # same negatives
    plain = ref.HSTU.fit(model, data)["rec_loss"]
    assert torch.allclose(fused, plain, rtol=1e-5)
    fused.backward()
    assert model.Item.embeddings.weight.grad is not None


@pytest.mark.parametrize("name,cls", [("MF-BPR", "MF"), ("LightGCN", "LightGCN")])
def test_genrec_pool(oracle_ops, name, cls):
    ref = shim.load_reference(name, embedding_dim=32)
    Base = getattr(ref, cls)
    Fused = type(cls + "B200", (arch.GenRecFused, Base), {})
    torch.manual_seed(0)
    model = Fused(shim.RecDataSet(n_users=40, n_items=150))
    model.eval()
    g = torch.Generator().manual_seed(6)
    data = {model.User: torch.arange(0, 40, 3).unsqueeze(1), model.IUnseen: torch.randint(0, 150, (14, 101), generator=g)}
    with torch.no_grad():
        model.reset_ranking_buffers()
        assert torch.allclose(model(data, ranking="pool"), Base.recommend_from_pool(model, data), rtol=1e-6, atol=1e-7)


def test_lightgcn_propagation(oracle_ops):
    """the mixin's encode (SpMM through ops.spmm) equals the reference's encode."""
    ref = shim.load_reference("LightGCN", embedding_dim=32)
    Fused = type("LightGCNB200", (arch.LightGCNFused, ref.LightGCN), {})
    torch.manual_seed(0)
    model = Fused(shim.RecDataSet(n_users=40, n_items=150))
    with torch.no_grad():
        model.User.embeddings.weight.normal_(0, 0.3)
        model.Item.embeddings.weight.normal_(0, 0.3)
    a, b = model.encode(), ref.LightGCN.encode(model)
    assert torch.allclose(a[0], b[0], rtol=1e-6, atol=1e-7) and torch.allclose(a[1], b[1], rtol=1e-6, atol=1e-7)
