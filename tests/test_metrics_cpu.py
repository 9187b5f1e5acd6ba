"""Host logic: metric reduction from a top-K list == the oracle's dense evaluation; synthetic
generators; shard bounds."""
import torch

from oracle import reference_path as orc
from recboard_b200 import metrics as MX
from recboard_b200 import sharded, synth

MONS = ["HITRATE@1", "HITRATE@5", "HITRATE@10", "NDCG@5", "NDCG@10", "RECALL@10", "PRECISION@5", "MRR@10"]


def _batch_metrics(ids, tcrow, tcol, n_items, mons):
    """The product's host-side reduction (metrics_from_hits) fed with the oracle's hit matrix: the product's own
    hit matrix is a kernel (rb_topk_hits) and exists on a GPU only."""
    n_t = (tcrow[1:] - tcrow[:-1]).float()
    return MX.metrics_from_hits(orc.hits_from_topk(ids, tcrow, tcol, n_items), n_t, mons)


def test_hit_matrix_has_no_cpu_path():
    import pytest
    ids = torch.zeros(2, 3, dtype=torch.int32)
    tcrow, tcol = MX.lists_to_csr([[0], [1]])
    with pytest.raises(RuntimeError):
        MX.hits_from_topk(ids, tcrow, tcol, 5)


def _case(seed, B=33, N=257, multi=False):
    g = torch.Generator().manual_seed(seed)
    scores = torch.randn(B, N, generator=g)
    seen = [torch.randint(0, N, (9,), generator=g).tolist() for _ in range(B)]
    tg = [torch.randint(0, N, (3 if multi else 1,), generator=g).tolist() for _ in range(B)]
    return scores, orc.lists_to_csr(seen), orc.lists_to_csr(tg)


def test_metrics_from_topk_bit_identical_to_dense_oracle():
    for seed, multi in ((0, False), (1, False), (2, True)):
        scores, (crow, col), (tcrow, tcol) = _case(seed, multi=multi)
        dense = orc.evaluate_batch(scores, crow, col, tcrow, tcol, MONS)
        _, ids = orc.topk_sorted(orc.mask_seen(scores, crow, col), 10)
        got = _batch_metrics(ids.int(), tcrow, tcol, scores.shape[1], MONS)
        assert got == dense  # bit-identical floats


def test_missing_entries_never_hit():
    ids = torch.tensor([[3, -1, -1]], dtype=torch.int32)
    tcrow, tcol = MX.lists_to_csr([[0]])
    assert _batch_metrics(ids, tcrow, tcol, 5, ["HITRATE@3"])["HITRATE@3"] == 0.0


def test_average_meter_weighting():
    m, o = MX.AverageMeter(), orc.AverageMeter()
    for v, n in ((0.25, 512), (0.5, 100)):
        m.update(v, n); o.update(v, n)
    assert m.avg == o.avg == (0.25 * 512 + 0.5 * 100) / 612


def test_shard_bounds_cover_catalog():
    for n, w in ((10, 3), (1_000_000, 8), (7, 8)):
        spans = [sharded.shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_merge_rowstats_matches_oracle():
    g = torch.Generator().manual_seed(5)
    U, W = torch.randn(20, 16, generator=g), torch.randn(90, 16, generator=g)
    lab = torch.randint(0, 90, (20,), generator=g)
    parts = []
    for a, b in ((0, 31), (31, 90)):
        S = orc.score_dense(U, W[a:b])
        m = S.max(1).values
        inside = (lab >= a) & (lab < b)
        ll = torch.where(inside, S.gather(1, (lab - a).clamp(0, b - a - 1)[:, None]).squeeze(1), torch.zeros(20))
        parts.append(torch.stack([m, torch.exp(S - m[:, None]).sum(1), ll]))
    lse, ll = sharded.merge_rowstats(torch.stack(parts))
    m, l, ll_ref = orc.ce_rowstats(U, W, lab)
    torch.testing.assert_close(lse, m + torch.log(l), rtol=1e-6, atol=1e-6)
    assert torch.equal(ll, ll_ref)


def test_synth_shapes_and_invariants():
    g = torch.Generator().manual_seed(0)
    dev = torch.device("cpu")
    ids = synth.zipf_ids(5000, 1000, g, dev)
    assert ids.min() >= 0 and ids.max() < 1000
    crow, col = synth.seen_csr(64, 1000, g, dev)
    assert crow[0] == 0 and crow[-1] == col.numel()
    for r in range(64):
        seg = col[crow[r]:crow[r + 1]]
        assert torch.all(seg[1:] > seg[:-1])  # sorted unique per row
    t = synth.targets(64, 1000, g, dev, (crow, col), frac_in_seen=1.0)
    assert all(int(t[r]) in col[crow[r]:crow[r + 1]].tolist() for r in range(64))
    s = synth.sequences(16, 50, 1000, g, dev)
    assert s.shape == (16, 50) and (s[:, -1] > 0).all() and s.max() <= 1000


def test_device_eval_split_matches_per_batch_csr(monkeypatch):
    """SURVEY 8f-4 host logic: the split-wide CSR cut per batch equals the CSR rebuilt from that batch's lists,
    and the sweep over it reproduces the oracle's bsz-weighted metrics (the hit-matrix kernel is a GPU-only
    product path: the oracle's dense-target gather stands in for it here)."""
    from recboard_b200 import evaluate as EV
    monkeypatch.setattr(MX, "hits_from_topk", orc.hits_from_topk)
    g = torch.Generator().manual_seed(11)
    R, N, K = 37, 200, 10
    seen = [torch.randperm(N, generator=g)[: int(torch.randint(0, 9, (1,), generator=g))].tolist() for _ in range(R)]
    tgt = [[int(torch.randint(0, N, (1,), generator=g))] for _ in range(R)]
    split = EV.DeviceEvalSplit(seen, tgt, torch.device("cpu"))
    for lo, hi in ((0, 16), (16, 32), (32, 37), (5, 5)):
        sc, sl, tc, tl = split.batch(lo, hi)
        rc, rl = orc.lists_to_csr(seen[lo:hi])
        assert torch.equal(sc, rc) and torch.equal(sl, rl)
        rc, rl = orc.lists_to_csr(tgt[lo:hi])
        assert torch.equal(tc, rc) and torch.equal(tl, rl)
    U, W = torch.randn(R, 16, generator=g), torch.randn(N, 16, generator=g)
    mons = ["HITRATE@5", "NDCG@10", "HITRATE@10"]

    def score_topk(lo, hi, k, crow, col):
        S = orc.score_dense(U[lo:hi], W)
        return orc.topk_sorted(orc.mask_seen(S, crow, col) if crow is not None else S, k)

    got = EV.evaluate_split(score_topk, split, mons, N, batch_size=16)
    meters = {m: MX.AverageMeter() for m in mons}
    for lo in range(0, R, 16):
        hi = min(lo + 16, R)
        sc, sl = orc.lists_to_csr(seen[lo:hi]); tc, tl = orc.lists_to_csr(tgt[lo:hi])
        for m, v in orc.evaluate_batch(orc.score_dense(U[lo:hi], W), sc, sl, tc, tl, mons).items():
            meters[m].update(v, hi - lo)
    assert got == {m: meters[m].avg for m in mons}


def test_device_eval_split_from_reference_rows():
    """8f-4, loader-facing half: rows in the format the reference's samplers yield (HSTU/sampler.py:107-125) become the
    finished batch tensors of the reference's pipes: last ``maxlen`` items, ids + NUM_PADS, left-padded with 0."""
    from recboard_b200 import evaluate as EV
    rows = [
        {"U": 3, "S": (5, 9, 2), "UNSEEN": (7,), "SEEN": (5, 9, 2)},
        {"U": 0, "S": (1, 1, 4, 8, 6, 3), "UNSEEN": (2,), "SEEN": (1, 4, 8, 6, 3)},
        {"U": 9, "S": (), "UNSEEN": (0, 4), "SEEN": ()},
    ]
    split = EV.DeviceEvalSplit.from_rows(rows, "U", "S", "UNSEEN", "SEEN", torch.device("cpu"), maxlen=4)
    assert split.users.tolist() == [[3], [0], [9]]
    assert split.seqs.tolist() == [[0, 6, 10, 3], [5, 9, 7, 4], [0, 0, 0, 0]]
    sc, sl, tc, tl = split.batch(0, 3)
    assert sc.tolist() == [0, 3, 8, 8] and sl.tolist() == [2, 5, 9, 1, 3, 4, 6, 8]
    assert tc.tolist() == [0, 1, 2, 4] and tl.tolist() == [7, 2, 0, 4]
    model = type("M", (), {"User": "User", "ISeq": "ISeq"})()
    d = split.data(1, 3, model)
    assert d["User"].tolist() == [[0], [9]] and d["ISeq"].shape == (2, 4)
