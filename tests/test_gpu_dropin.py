"""The drop-in layer on a GPU: every mixin of ``recboard_b200.arch`` behind models that honour the
reference's ``RecSysArch`` contract (tests/dropin_models.py -- the GPU box has neither /root/reference nor
freerec), the evaluation sweeps of ``recboard_b200.evaluate`` including ``FusedEvalCoach.evaluate`` behind a
stub ``Coach``, and a 20-step Adam loop whose loss curve must follow the eager (reference-lines) model's.

The eager side of every comparison is the model's own plain-PyTorch statement of the reference lines
(SASRec/main.py:195-236, UniSRec/main.py:400-447) in fp32 on the same device; evaluation metrics are checked
against the CPU oracle."""
import copy
import types

import pytest
import torch

from oracle import reference_path as orc
from recboard_b200 import arch, evaluate as EV, metrics as MX
import dropin_models as DM   # tests/ is on sys.path (pytest rootdir-less import mode; conftest adds the repo root)

pytestmark = pytest.mark.gpu
FP32_RTOL = 1e-5


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    return torch.device("cuda", 0)


def fuse(mixin, base, **attrs):
    return type(base.__name__ + "B200", (mixin, base), attrs)


def rel(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def grads_of(model):
    return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}


def _fit_pair(Fused, Base, data_fn, cuda, seed=0, **kw):
    """Same weights, same batch: fused fit vs the eager reference lines -> (loss_f, loss_e, grads_f, grads_e)."""
    torch.manual_seed(seed)
    eager = Base(**kw).to(cuda).train()
    fused = Fused(**kw).to(cuda).train()
    fused.load_state_dict(eager.state_dict())
    out = []
    for model in (fused, eager):
        torch.manual_seed(123)   # same random mask / negatives on both sides
        loss = model(data_fn(model))["rec_loss"]
        loss.backward()
        out.append((loss.detach(), grads_of(model)))
    return fused, eager, out[0], out[1]


def _seq_data(cuda, B=9, S=12, N=300, pads=1, seed=1):
    g = torch.Generator().manual_seed(seed)
    seqs = DM.lpad_sequences(g, B, S, N, pads, cuda)
    pos = torch.randint(0, N, (B, S), generator=g).to(cuda)
    pool = torch.randint(0, N, (B, 17), generator=g).to(cuda)
    return lambda m: {m.ISeq: seqs.clone(), m.IPos: pos, m.IUnseen: pool}


def _check_eval(fused, eager, data_fn, N, K=10, tol=FP32_RTOL):
    fused.eval(); eager.eval()
    fused.reset_ranking_buffers(); eager.reset_ranking_buffers()
    with torch.no_grad():
        full_f, full_e = fused(data_fn(fused), ranking="full"), eager(data_fn(eager), ranking="full")
        assert full_f.shape == full_e.shape and full_f.shape[1] == N
        assert rel(full_f, full_e) <= tol
        pool_f, pool_e = fused(data_fn(fused), ranking="pool"), eager(data_fn(eager), ranking="pool")
        assert rel(pool_f, pool_e) <= tol
        B = full_e.shape[0]
        g = torch.Generator().manual_seed(5)
        seen = [torch.randperm(N, generator=g)[:7].tolist() for _ in range(B)]
        crow, col = orc.lists_to_csr(seen)
        vals, ids = fused.recommend_topk(data_fn(fused), K, crow.cuda(), col.cuda())
        masked = orc.mask_seen(full_e.cpu().clone(), crow, col)        # scores[seen] = -1e23, UniSRec/main.py:409-413
        rv, ri = orc.topk_sorted(masked, K)
        assert rel(vals, rv) <= tol
        gaps_ok = (rv[:, :-1] - rv[:, 1:]).abs() > 4 * tol * rv.abs().max()
        same = ids.cpu().long() == ri
        assert bool((same[:, :-1] | ~gaps_ok).all())   # identical ids wherever the reference's neighbours are separated


# ------------------------------------------------------------------------------------- mixins
def test_sasrec_fused_fit_full_pool_topk(cuda):
    N = 300
    F_ = fuse(arch.SASRecFused, DM.TinySASRec)
    data_fn = _seq_data(cuda, N=N)
    fused, eager, (lf, gf), (le, ge) = _fit_pair(F_, DM.TinySASRec, data_fn, cuda, n_users=9, n_items=N)
    assert abs(float(lf) - float(le)) <= FP32_RTOL * abs(float(le))
    assert set(gf) == set(ge)
    for name in ge:
        assert rel(gf[name], ge[name]) <= 2e-5, name       # fp32-parity mode end to end, encoder included
    _check_eval(fused, eager, data_fn, N)


def test_sasrec_fused_sync_free_fit(cuda):
    """fused_sync_free: the non-padding query rows are compacted on the device; ``fit`` runs without any device->host
    synchronisation and reproduces the eager loss and gradients."""
    N = 300
    F_ = fuse(arch.SASRecFused, DM.TinySASRec, fused_sync_free=True)
    data_fn = _seq_data(cuda, N=N)
    torch.manual_seed(0)
    eager = DM.TinySASRec(n_users=9, n_items=N).to(cuda).train()
    fused = F_(n_users=9, n_items=N).to(cuda).train()
    fused.load_state_dict(eager.state_dict())
    le = eager(data_fn(eager))["rec_loss"]
    le.backward()
    data = data_fn(fused)
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        lf = fused(data)["rec_loss"]
        lf.backward()
    finally:
        torch.cuda.set_sync_debug_mode("default")
    assert abs(float(lf) - float(le)) <= FP32_RTOL * abs(float(le))
    ge, gf = grads_of(eager), grads_of(fused)
    for name in ge:
        assert rel(gf[name], ge[name]) <= 2e-5, name


def test_sasrec_fused_bf16_precision(cuda):
    N = 300
    F_ = fuse(arch.SASRecFused, DM.TinySASRec, fused_precision="bf16")
    data_fn = _seq_data(cuda, N=N)
    fused, eager, (lf, gf), (le, ge) = _fit_pair(F_, DM.TinySASRec, data_fn, cuda, n_users=9, n_items=N)
    assert abs(float(lf) - float(le)) <= 5e-3 * abs(float(le))       # operands rounded to bf16 inside the fused op
    for name in ge:
        assert rel(gf[name], ge[name]) <= 3e-2, name


def test_gru4rec_fused(cuda):
    N = 250
    g = torch.Generator().manual_seed(2)
    seqs = DM.lpad_sequences(g, 8, 12, N, 1, cuda)
    pos = torch.randint(0, N, (8, 1), generator=g).to(cuda)
    pool = torch.randint(0, N, (8, 11), generator=g).to(cuda)
    data_fn = lambda m: {m.ISeq: seqs, m.IPos: pos, m.IUnseen: pool}
    fused, eager, (lf, gf), (le, ge) = _fit_pair(fuse(arch.GRU4RecFused, DM.TinyGRU4Rec), DM.TinyGRU4Rec, data_fn, cuda,
                                                 n_users=8, n_items=N)
    assert abs(float(lf) - float(le)) <= FP32_RTOL * abs(float(le))
    for name in ge:
        assert rel(gf[name], ge[name]) <= 5e-5, name
    _check_eval(fused, eager, data_fn, N)


def test_bert4rec_fused_bias_head(cuda):
    N = 200
    data_fn = _seq_data(cuda, B=7, N=N, pads=2, seed=3)
    fused, eager, (lf, gf), (le, ge) = _fit_pair(fuse(arch.BERT4RecFused, DM.TinyBERT4Rec), DM.TinyBERT4Rec, data_fn, cuda,
                                                 n_users=7, n_items=N)
    assert abs(float(lf) - float(le)) <= FP32_RTOL * abs(float(le))
    assert "fc.bias" in gf and "fc.weight" in gf
    for name in ge:
        assert rel(gf[name], ge[name]) <= 5e-5, name
    _check_eval(fused, eager, data_fn, N)


def test_hstu_fused_sampled_fit_and_cached_table(cuda):
    N = 280
    F_ = fuse(arch.HSTUFused, DM.TinyHSTU)
    data_fn = _seq_data(cuda, N=N, seed=4)
    fused, eager, (lf, gf), (le, ge) = _fit_pair(F_, DM.TinyHSTU, data_fn, cuda, n_users=9, n_items=N)
    assert abs(float(lf) - float(le)) <= FP32_RTOL * abs(float(le))
    for name in ge:
        assert rel(gf[name], ge[name]) <= 5e-5, name
    _check_eval(fused, eager, data_fn, N)     # reset_ranking_buffers -> cached normalised table
    assert fused._fused_item is not None
    # the cache must not survive a weight update (stale embeddings after more training)
    with torch.no_grad():
        fused.Item.embeddings.weight.add_(0.05 * torch.randn_like(fused.Item.embeddings.weight))
        eager.load_state_dict(fused.state_dict())
        assert rel(fused(data_fn(fused), ranking="full"), eager(data_fn(eager), ranking="full")) <= FP32_RTOL


@pytest.mark.parametrize("base,mixin", [(DM.TinyMF, arch.GenRecFused), (DM.TinyLightGCN, arch.LightGCNFused)])
def test_genrec_fused(cuda, base, mixin):
    U_, N = 40, 260
    g = torch.Generator().manual_seed(6)
    users = torch.arange(0, U_, 3).unsqueeze(1).to(cuda)
    B = users.shape[0]
    pos = torch.randint(0, N, (B, 1), generator=g).to(cuda)
    neg = torch.randint(0, N, (B, 1), generator=g).to(cuda)
    pool = torch.randint(0, N, (B, 13), generator=g).to(cuda)
    data_fn = lambda m: {m.User: users, m.IPos: pos, m.INeg: neg, m.IUnseen: pool}
    fused, eager, (lf, gf), (le, ge) = _fit_pair(fuse(mixin, base), base, data_fn, cuda, n_users=U_, n_items=N)
    assert abs(float(lf) - float(le)) <= FP32_RTOL * abs(float(le))   # BPR fit stays the reference's; LightGCN's encode is ops.spmm
    for name in ge:
        assert rel(gf[name], ge[name]) <= 5e-5, name
    _check_eval(fused, eager, data_fn, N)


def test_fused_embedding_swaps_in_and_trains(cuda):
    """``arch.fuse_item_embedding``: the lookup and its dense backward through the C ABI, same weight Parameter;
    with ``accumulate_grad`` the rows land in the existing ``weight.grad`` next to the head's dW."""
    N = 300
    data_fn = _seq_data(cuda, N=N)
    F_ = fuse(arch.SASRecFused, DM.TinySASRec)
    fused, eager, (lf, gf), (le, ge) = _fit_pair(F_, DM.TinySASRec, data_fn, cuda, n_users=9, n_items=N)
    for accumulate in (False, True):
        m2 = F_(n_users=9, n_items=N).to(cuda).train()
        m2.load_state_dict(eager.state_dict())
        arch.fuse_item_embedding(m2, accumulate_grad=accumulate)
        assert set(m2.state_dict()) == set(eager.state_dict())
        if accumulate:
            for p in m2.parameters():
                p.grad = torch.zeros_like(p)
        torch.manual_seed(123)
        loss = m2(data_fn(m2))["rec_loss"]
        loss.backward()
        assert abs(float(loss) - float(le)) <= FP32_RTOL * abs(float(le))
        for name, p in m2.named_parameters():
            assert rel(p.grad, ge[name]) <= 2e-5, (name, accumulate)


# ------------------------------------------------------------------------------ evaluation sweeps
MONS = ["HITRATE@1", "HITRATE@5", "HITRATE@10", "NDCG@5", "NDCG@10"]


def _eval_setup(cuda, R=70, N=300, bs=32):
    torch.manual_seed(7)
    model = fuse(arch.SASRecFused, DM.TinySASRec)(n_users=R, n_items=N).to(cuda).eval()
    g = torch.Generator().manual_seed(8)
    seqs = DM.lpad_sequences(g, R, 12, N, 1, "cpu")
    seen = [sorted(set((seqs[r][seqs[r] > 0] - 1).tolist())) for r in range(R)]   # ISeen = the items of the sequence
    tgt = [[int(torch.randint(0, N, (1,), generator=g))] for _ in range(R)]
    tgt[3] = [seen[3][0]]                                                         # a target inside the seen list never hits
    batches = []
    for lo in range(0, R, bs):
        hi = min(lo + bs, R)
        batches.append({model.ISeq: seqs[lo:hi], model.ISeen: seen[lo:hi], model.IUnseen: tgt[lo:hi], model.Size: hi - lo})
    return model, seqs, seen, tgt, batches


def _oracle_sweep(model, seqs, seen, tgt, bs, cuda):
    rows = []
    with torch.no_grad():
        for lo in range(0, len(seen), bs):
            hi = min(lo + bs, len(seen))
            S = DM.TinySASRec.recommend_from_full(model, {model.ISeq: seqs[lo:hi].to(cuda)}).cpu()
            rows.append((S, *orc.lists_to_csr(seen[lo:hi]), *orc.lists_to_csr(tgt[lo:hi])))
    return orc.evaluate_sweep(rows, MONS)                                         # UniSRec/main.py:400-447 restated


def test_evaluate_sweep_and_split_match_oracle(cuda):
    model, seqs, seen, tgt, batches = _eval_setup(cuda)
    ref = _oracle_sweep(model, seqs, seen, tgt, 32, cuda)
    dev_batches = [{k: (v.to(cuda) if isinstance(v, torch.Tensor) else v) for k, v in b.items()} for b in batches]
    got = EV.evaluate_sweep(model, dev_batches, MONS, model.Item.count, size_key=model.Size)
    assert got == ref
    split = EV.DeviceEvalSplit(seen, tgt, cuda)                                   # seen / target CSR on the device once
    seqs_d = seqs.to(cuda)

    def score_topk(lo, hi, k, crow, col):
        return model.recommend_topk({model.ISeq: seqs_d[lo:hi]}, k, crow, col)

    assert EV.evaluate_split(score_topk, split, MONS, model.Item.count, batch_size=32) == ref
    # metrics reduced on the device in one pass per batch, one read-back per sweep: same values to 1e-6
    fast = EV.evaluate_split(score_topk, split, MONS, model.Item.count, batch_size=32, exact=False)
    assert set(fast) == set(ref) and all(abs(fast[k] - ref[k]) <= 1e-6 * max(1.0, abs(ref[k])) for k in ref)


def test_evaluate_split_model_from_reference_rows(cuda):
    """8f-4 end to end on the GPU: rows in the reference samplers' format -> DeviceEvalSplit.from_rows -> a whole
    evaluation sweep without per-batch host work, equal to the oracle's sweep over the same rows."""
    model, seqs, seen, tgt, batches = _eval_setup(cuda)
    rows = []
    for r in range(len(seen)):
        items = (seqs[r][seqs[r] > 0] - 1).tolist()                        # 0-based item ids, no padding
        rows.append({"U": r, "S": tuple(items), "UNSEEN": tuple(tgt[r]), "SEEN": tuple(seen[r])})
    split = EV.DeviceEvalSplit.from_rows(rows, "U", "S", "UNSEEN", "SEEN", cuda, maxlen=12)
    assert torch.equal(split.seqs.cpu(), seqs)                              # the pipes' add_(NUM_PADS) + lpad_ reproduced
    got = EV.evaluate_split_model(model, split, MONS, batch_size=32)
    assert got == _oracle_sweep(model, seqs, seen, tgt, 32, cuda)


class _StubCoach:
    """What ``FusedEvalCoach`` needs from freerec's ``Coach`` (UniSRec/main.py:400-447): cfg, fields, dataloader,
    dict_to_device, register_metric and a ``monitor`` that keeps bsz-weighted means."""

    def __init__(self, model, batches, device):
        self.cfg = types.SimpleNamespace(monitors=["LOSS"] + MONS, ranking="full")
        self.model, self.dataloader, self.device, self.remove_seen = model, batches, device, True
        self.ISeen, self.IUnseen, self.Size = model.ISeen, model.IUnseen, model.Size
        self.meters, self.registered = {}, {}

    def get_res_sys_arch(self):
        return self.model

    def dict_to_device(self, data):
        return {k: (v.to(self.device) if isinstance(v, torch.Tensor) else v) for k, v in data.items()}

    def register_metric(self, name, func, fmt=".4f", best_caster=max):
        self.registered[name] = func

    def monitor(self, *values, n=1, reduction="mean", mode="valid", pool=None):
        for name in pool:
            self.meters.setdefault((mode, name), orc.AverageMeter()).update(self.registered[name](*values), n)


def test_fused_eval_coach_evaluate(cuda):
    model, seqs, seen, tgt, batches = _eval_setup(cuda)
    Coach = type("CoachB200", (EV.FusedEvalCoach, _StubCoach), {})
    coach = Coach(model, batches, cuda)
    coach.set_other()
    assert set(coach.registered) == set(MONS)
    coach.evaluate(epoch=0, mode="test")
    ref = _oracle_sweep(model, seqs, seen, tgt, 32, cuda)
    assert {k[1]: m.avg for k, m in coach.meters.items()} == ref


# ------------------------------------------------------------------------------ training loop
@pytest.mark.parametrize("precision,tol", [(None, 5e-4), ("bf16", 3e-2)])
def test_twenty_adam_steps_follow_eager(cuda, precision, tol):
    """Coach.train_per_epoch (SASRec/main.py:242-258): 20 optimizer steps, fused vs eager, same init and batches."""
    N, B, S = 500, 32, 12
    torch.manual_seed(9)
    eager = DM.TinySASRec(n_users=B, n_items=N).to(cuda).train()
    fused = fuse(arch.SASRecFused, DM.TinySASRec, fused_precision=precision)(n_users=B, n_items=N).to(cuda).train()
    fused.load_state_dict(copy.deepcopy(eager.state_dict()))
    opts = [torch.optim.Adam(m.parameters(), lr=1e-3) for m in (fused, eager)]
    g = torch.Generator().manual_seed(10)
    curves = ([], [])
    for step in range(20):
        seqs = DM.lpad_sequences(g, B, S, N, 1, cuda)
        pos = torch.randint(0, N, (B, S), generator=g).to(cuda)
        for model, opt, curve in zip((fused, eager), opts, curves):
            loss = model({model.ISeq: seqs, model.IPos: pos})["rec_loss"]      # :247-248
            opt.zero_grad()
            loss.backward()                                                    # :249
            opt.step()
            curve.append(float(loss))
    f, e = torch.tensor(curves[0]), torch.tensor(curves[1])
    assert float(((f - e).abs() / e.abs()).max()) <= tol, (curves[0], curves[1])
    assert e[-1] < e[0] and f[-1] < f[0]
