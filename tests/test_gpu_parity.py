"""Parity tests proper: the CUDA path (through the C ABI) vs the CPU oracle on the same seeded
inputs and against the golden vectors produced by the reference's own model files.

Tolerances (BASELINE.json north_star): fp32 mode 1e-5 relative; bf16 mode 2e-3 relative (the bf16
inputs are shared with the oracle, so forward values agree far tighter; gradients carry one bf16
rounding of the softmax tile).  Top-K ids must be identical wherever neighbouring reference scores
are further apart than the fp32 tolerance.
"""
import numpy as np
import pytest
import torch

from oracle import reference_path as orc
from recboard_b200 import metrics as MX

pytestmark = pytest.mark.gpu

FP32_RTOL = 1e-5
BF16_RTOL = 2e-3
T = torch.from_numpy


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    from recboard_b200 import ops as _ops
    return _ops


def dev(x):
    return None if x is None else x.cuda()


def assert_rel(got, ref, rtol, what=""):
    got, ref = got.detach().float().cpu(), torch.as_tensor(ref).float()
    scale = float(ref.abs().max().clamp_min(1e-30))
    err = float((got - ref).abs().max()) / scale
    assert err <= rtol, f"{what}: max err / max|ref| = {err:.3e} > {rtol:.1e}"


def assert_grad_bf16(got, ref, what=""):
    """bf16-mode gradients against the fp32 oracle, north_star's 2e-3 in two norms: the largest error relative to the
    largest element (plus half a bf16 ulp of it, 2^-9, when the gradient itself is STORED in bf16 -- a bf16
    parameter's gradient), and the Euclidean norm of the error relative to that of the reference.  (An element-by-
    element relative bound is not meaningful here: the one bf16 rounding of the softmax tile gives every element an
    ABSOLUTE error proportional to the size of its summands, and dU = sum_j P_ij w_j - w_label cancels by design.)"""
    stored_bf16 = got.dtype == torch.bfloat16
    got, ref = got.detach().float().cpu(), torch.as_tensor(ref).float()
    mx = float(ref.abs().max().clamp_min(1e-30))
    tol = BF16_RTOL + (2.0 ** -9 if stored_bf16 else 0.0)
    err = float((got - ref).abs().max()) / mx
    assert err <= tol, f"{what}: max err / max|ref| = {err:.3e} > {tol:.1e}"
    l2 = float((got - ref).double().norm() / ref.double().norm().clamp_min(1e-30))
    assert l2 <= tol, f"{what}: |err|_2 / |ref|_2 = {l2:.3e} > {tol:.1e}"


def bf16_round(x):
    return x.bfloat16().float()


# ------------------------------------------------------------------------------- golden vectors
def test_golden_sasrec_ce_fp32_forward(ops, golden):
    g = golden("sasrec_ce")
    U, W, lab = T(g["U"]), T(g["W"]), T(g["labels"])
    m, l, ll = ops.ce_rowstats(dev(U), dev(W), dev(lab), precision="fp32")
    loss = (m + torch.log(l) - ll).mean()
    assert abs(float(loss) - float(g["loss"])) <= FP32_RTOL * abs(float(g["loss"]))
    assert_rel(ops.score_dense(dev(T(g["U_eval"])), dev(W), precision="fp32"), g["scores_full"], FP32_RTOL, "scores")


def test_golden_sasrec_ce_bf16_loss_and_grads(ops, golden):
    g = golden("sasrec_ce")
    U, W, lab = bf16_round(T(g["U"])), bf16_round(T(g["W"])), T(g["labels"])
    ref_loss, ref_dU, ref_dW, _ = orc.ce_fwd_bwd(U, W, lab)
    Ud, Wd = dev(U).bfloat16().requires_grad_(True), dev(W).bfloat16().requires_grad_(True)
    loss = ops.fused_ce(Ud, Wd, dev(lab))
    loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    assert_grad_bf16(Ud.grad, ref_dU, "dU")
    assert_grad_bf16(Wd.grad, ref_dW, "dW")
    # and against the reference's own numbers (fp32 inputs): within the bf16 tolerance
    assert abs(float(loss) - float(g["loss"])) <= BF16_RTOL * abs(float(g["loss"]))


def test_golden_bert4rec_bias_head(ops, golden):
    g = golden("bert4rec_ce")
    U, W, b, lab = T(g["U"]), T(g["W"]), T(g["bias"]), T(g["labels"])
    m, l, ll = ops.ce_rowstats(dev(U), dev(W), dev(lab), bias=dev(b), precision="fp32")
    loss = (m + torch.log(l) - ll).mean()
    assert abs(float(loss) - float(g["loss"])) <= FP32_RTOL * abs(float(g["loss"]))
    Ub, Wb = bf16_round(U), bf16_round(W)
    _, rdU, rdW, rdb = orc.ce_fwd_bwd(Ub, Wb, lab, bias=b)
    lse = torch.logsumexp(orc.score_dense(Ub, Wb, b), dim=1)
    dU, dW, db = ops.ce_backward(dev(Ub).bfloat16(), dev(Wb).bfloat16(), dev(lab), dev(lse), 1.0 / len(lab), bias=dev(b),
                                 need_dbias=True)
    assert_grad_bf16(dU, rdU, "dU")
    assert_grad_bf16(dW, rdW, "dW")
    assert_grad_bf16(db, rdb, "dbias")
    f = golden("bert4rec_full")
    S = ops.score_dense(dev(T(f["U"])), dev(T(f["W"])), bias=dev(T(f["bias"])), precision="fp32")[:, int(f["num_pads"]):]
    assert_rel(S, f["scores_full"], FP32_RTOL, "bert4rec scores")


@pytest.mark.parametrize("name", ["mf_full", "lightgcn_full"])
def test_golden_genrec_full_ranking(ops, golden, name):
    g = golden(name)
    users = T(g["users"]).squeeze(1)
    U = ops.gather_rows_raw(dev(T(g["user_table"])), dev(users))
    assert torch.equal(U.cpu(), T(g["user_table"])[users])
    assert_rel(ops.score_dense(U, dev(T(g["item_table"])), precision="fp32"), g["scores_full"], FP32_RTOL, name)


def test_golden_hstu_and_gru4rec_scores(ops, golden):
    for name in ("hstu_full", "gru4rec_ce"):
        g = golden(name)
        Uk = "U" if name == "hstu_full" else "U_eval"
        assert_rel(ops.score_dense(dev(T(g[Uk])), dev(T(g["W"])), precision="fp32"), g["scores_full"], FP32_RTOL, name)


def test_golden_pool_sampled_and_propagation(ops, golden):
    """The f-rows against the numbers the reference's own model files produced: SASRec / HSTU pool scores,
    HSTU table normalisation + sampled-softmax fit (loss, dU, dense dW), LightGCN propagation."""
    g = golden("sasrec_pool")
    assert_rel(ops.gather_dot(dev(T(g["U"])), dev(T(g["W"])), dev(T(g["pool"]))), g["scores_pool"], FP32_RTOL, "sasrec pool")
    h = golden("hstu_sampled")
    Wn = ops.normalize_rows(dev(T(h["table"]))[1:])
    assert_rel(Wn, h["W_norm"], 1e-6, "normalised table")
    assert_rel(ops.gather_dot(dev(T(h["U_pool"])), Wn, dev(T(h["pool"]))), h["scores_pool"], FP32_RTOL, "hstu pool")
    U, W = dev(T(h["U_fit"])).requires_grad_(True), dev(T(h["W_fit"])).requires_grad_(True)
    cand = torch.cat((T(h["positives"]).unsqueeze(-1), T(h["negatives"])), dim=1)
    logits = ops.gather_dot(U, W, dev(cand), scale=1.0 / float(h["temperature"]))
    loss = torch.nn.functional.cross_entropy(logits, torch.zeros(len(U), dtype=torch.long, device="cuda"))
    loss.backward()
    assert abs(float(loss.detach()) - float(h["loss"])) <= FP32_RTOL * abs(float(h["loss"]))
    assert_rel(U.grad, h["dU_fit"], 2e-5, "sampled-softmax dU")
    assert_rel(W.grad, h["dW_fit"], 2e-5, "sampled-softmax dW")
    p = golden("lightgcn_prop")
    nU = p["user_table"].shape[0]
    n = nU + p["item_table"].shape[0]
    A = torch.sparse_csr_tensor(T(p["crow"]), T(p["col"]), T(p["val"]), (n, n)).cuda()
    L = int(p["num_layers"])
    x = dev(torch.cat((T(p["user_table"]), T(p["item_table"]))))
    avg = x / (L + 1)
    for _ in range(L):
        x = ops.spmm_raw(A, x, acc=avg, beta=1.0 / (L + 1))    # propagation + layer average in one pass
    assert_rel(avg[:nU], p["user_out"], FP32_RTOL, "lightgcn users")
    assert_rel(avg[nU:], p["item_out"], FP32_RTOL, "lightgcn items")


def test_golden_embedding_forward_backward(ops, golden):
    g = golden("embedding_bwd")
    table, idx, go = T(g["table"]), T(g["idx"]), T(g["grad_out"])
    t = dev(table).requires_grad_(True)
    out = ops.gather_rows(t, dev(idx), padding_idx=0)
    assert torch.equal(out.detach().cpu(), T(g["out"]))
    out.backward(dev(go))
    assert_rel(t.grad, g["grad_table"], 1e-6, "embedding backward")
    assert torch.all(t.grad[0] == 0)  # padding_idx row (SASRec/main.py:75)


# ----------------------------------------------------------------------- seeded oracle parity
SHAPES = [  # (M, N, d) incl. partial tiles, M=1, tiny N, label in last partial tile
    (1, 130, 64), (77, 129, 32), (300, 1000, 64), (513, 4099, 128), (256, 2048, 256), (129, 257, 8),
]


@pytest.mark.parametrize("M,N,d", SHAPES)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_scores_and_rowstats(ops, M, N, d, precision):
    if precision == "fp32" and d > 128:
        pytest.skip("fp32x3 mode supports d <= 128")
    g = torch.Generator().manual_seed(M * 7 + N)
    U = torch.randn(M, d, generator=g) / d ** 0.25
    W = torch.randn(N, d, generator=g) / d ** 0.25
    if precision == "bf16":
        U, W = bf16_round(U), bf16_round(W)
    lab = torch.randint(0, N, (M,), generator=g)
    lab[-1] = N - 1  # label in the last (partial) tile
    bias = torch.randn(N, generator=g) * 0.3
    rtol = FP32_RTOL if precision == "fp32" else 1e-5
    cast = (lambda x: dev(x).bfloat16()) if precision == "bf16" else dev
    for b, scale in ((None, 1.0), (bias, 0.7)):
        S = ops.score_dense(cast(U), cast(W), bias=dev(b), scale=scale, precision=precision)
        assert_rel(S, orc.score_dense(U, W, b, scale), rtol, "scores")
        m, l, ll = ops.ce_rowstats(cast(U), cast(W), dev(lab), bias=dev(b), scale=scale, precision=precision)
        rm, rl, rll = orc.ce_rowstats(U, W, lab, b, scale)
        lse, rlse = (m + torch.log(l)).cpu(), rm + torch.log(rl)
        assert float((lse - rlse).abs().max()) <= rtol * float(rlse.abs().max()) + 1e-6
        assert_rel(ll, rll, rtol, "label logit")


@pytest.mark.parametrize("M,N,d", [(1, 130, 64), (300, 1000, 64), (513, 4099, 128), (3013, 12101, 64),
                                   (17000, 700, 64),    # > 16384 rows: one-hot correction through the sorted scatter
                                   (300, 1000, 256), (513, 4099, 192), (2100, 40_000, 256),   # 128 < d <= 256: d-split passes
                                   (5000, 1000, 256)])    # ... with the query range of the dW pass cut into several splits
def test_ce_gradients_bf16(ops, M, N, d):
    g = torch.Generator().manual_seed(M + N + d)
    U = bf16_round(torch.randn(M, d, generator=g) * 1.5 / d ** 0.25)
    W = bf16_round(torch.randn(N, d, generator=g) * 1.5 / d ** 0.25)
    lab = torch.randint(0, N, (M,), generator=g)
    lab[: max(1, M // 8)] = 3  # hot label row: many queries share one item
    ref_loss, rdU, rdW, _ = orc.ce_fwd_bwd(U, W, lab, scale=0.9, grad_out=2.0)
    Ud, Wd = dev(U).bfloat16().requires_grad_(True), dev(W).bfloat16().requires_grad_(True)
    loss = ops.fused_ce(Ud, Wd, dev(lab), scale=0.9)
    (loss * 2.0).backward()
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    assert_grad_bf16(Ud.grad, rdU, "dU")
    assert_grad_bf16(Wd.grad, rdW, "dW")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_embedding_width_not_a_multiple_of_8(ops, precision):
    """d = 50 (HSTU/configs/MovieLens1M_500_LOU.yaml): the host layer zero-pads the columns; scores, CE loss and
    gradients (true width), masked top-K, cosine normalisation and pool scores against the oracle."""
    g = torch.Generator().manual_seed(50)
    M, N, d, K = 130, 3706, 50, 20
    U, W = torch.randn(M, d, generator=g) * 0.4, torch.randn(N, d, generator=g) * 0.4
    if precision == "bf16":
        U, W = bf16_round(U), bf16_round(W)
    lab = torch.randint(0, N, (M,), generator=g)
    cast = (lambda x: dev(x).bfloat16()) if precision == "bf16" else dev
    rtol = FP32_RTOL if precision == "fp32" else 1e-5
    assert_rel(ops.score_dense(cast(U), cast(W), precision=precision), orc.score_dense(U, W), rtol, "scores")
    ref_loss, rdU, rdW, _ = orc.ce_fwd_bwd(U, W, lab)
    Ud, Wd = cast(U).requires_grad_(True), cast(W).requires_grad_(True)
    loss = ops.fused_ce(Ud, Wd, dev(lab), precision=precision)
    loss.backward()
    assert Ud.grad.shape == (M, d) and Wd.grad.shape == (N, d)
    assert abs(float(loss) - float(ref_loss)) <= rtol * abs(float(ref_loss))
    if precision == "fp32":
        assert_rel(Ud.grad, rdU, 2e-5, "dU"); assert_rel(Wd.grad, rdW, 2e-5, "dW")
    else:
        assert_grad_bf16(Ud.grad, rdU, "dU"); assert_grad_bf16(Wd.grad, rdW, "dW")
    seen = _seen_lists(g, M, N, 30)
    crow, col = MX.lists_to_csr(seen)
    vals, ids = ops.topk_eval(cast(U), cast(W), K, dev(crow), dev(col), precision=precision)
    _check_topk(vals, ids, orc.mask_seen(orc.score_dense(U, W), crow, col), K, 1e-5 if precision == "fp32" else 2e-5)
    Wn = ops.normalize_rows(dev(W))
    assert Wn.shape == (N, d)
    assert_rel(Wn, torch.nn.functional.normalize(W, dim=-1), 1e-6, "normalised rows")
    idx = torch.randint(0, N, (M, 7), generator=g)
    S = ops.gather_dot(cast(U), cast(W), dev(idx))
    assert_rel(S, torch.einsum("md,mkd->mk", U, W[idx]), rtol, "pool scores")


@pytest.mark.parametrize("M,N,d", [(700, 5000, 256), (257, 1300, 200)])
def test_ce_bias_head_wide_rows(ops, M, N, d):
    """128 < d <= 256 (CCFRec / E4SRec-style heads, config-4-shaped training): the d-split fused passes with bias and a
    scale, through autograd -- loss, dU, dW, dbias against the oracle."""
    g = torch.Generator().manual_seed(M + N + d)
    U = bf16_round(torch.randn(M, d, generator=g) * 1.2 / d ** 0.25)
    W = bf16_round(torch.randn(N, d, generator=g) * 1.2 / d ** 0.25)
    b = torch.randn(N, generator=g) * 0.3
    lab = torch.randint(0, N, (M,), generator=g)
    lab[: M // 6] = N - 1
    ref_loss, rdU, rdW, rdb = orc.ce_fwd_bwd(U, W, lab, bias=b, scale=0.8)
    Ud, Wd, bd = dev(U).bfloat16().requires_grad_(True), dev(W).bfloat16().requires_grad_(True), dev(b).requires_grad_(True)
    loss = ops.fused_ce(Ud, Wd, dev(lab), bias=bd, scale=0.8)
    loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    assert_grad_bf16(Ud.grad, rdU, "dU")
    assert_grad_bf16(Wd.grad, rdW, "dW")
    assert_grad_bf16(bd.grad, rdb, "dbias")


@pytest.mark.parametrize("M,N,d,with_bias", [(1, 130, 64, False), (300, 1000, 64, True), (513, 4099, 32, False),
                                               (3013, 12101, 64, False), (700, 5000, 128, True), (260, 3000, 96, False)])
def test_ce_gradients_fp32(ops, M, N, d, with_bias):
    """fp32-parity mode trains too: loss and all gradients within 1e-5 of the oracle (config 1 shape last)."""
    g = torch.Generator().manual_seed(7 * M + N + d)
    U = torch.randn(M, d, generator=g) * 1.5 / d ** 0.25
    W = torch.randn(N, d, generator=g) * 1.5 / d ** 0.25
    b = torch.randn(N, generator=g) * 0.3 if with_bias else None
    lab = torch.randint(0, N, (M,), generator=g)
    lab[: max(1, M // 8)] = 3
    ref_loss, rdU, rdW, rdb = orc.ce_fwd_bwd(U, W, lab, b, scale=0.9, grad_out=2.0)
    Ud, Wd = dev(U).requires_grad_(True), dev(W).requires_grad_(True)
    bd = dev(b).requires_grad_(True) if with_bias else None
    loss = ops.fused_ce(Ud, Wd, dev(lab), bias=bd, scale=0.9, precision="fp32")
    (loss * 2.0).backward()
    assert abs(float(loss.detach()) - float(ref_loss)) <= FP32_RTOL * abs(float(ref_loss))
    assert_rel(Ud.grad, rdU, FP32_RTOL, "dU")
    assert_rel(Wd.grad, rdW, FP32_RTOL, "dW")
    if with_bias:
        assert_rel(bd.grad, rdb, FP32_RTOL, "dbias")


def test_golden_ce_fp32_gradients(ops, golden):
    """Gradients in fp32-parity mode against the numbers the reference's own model files produced:
    SASRec dU (SASRec/main.py:217-219 + autograd) and the BERT4Rec bias head's dU / dW / dbias."""
    g = golden("sasrec_ce")
    Ud, Wd = dev(T(g["U"])).requires_grad_(True), dev(T(g["W"])).requires_grad_(True)
    loss = ops.fused_ce(Ud, Wd, dev(T(g["labels"])), precision="fp32")
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) <= FP32_RTOL * abs(float(g["loss"]))
    assert_rel(Ud.grad, g["dU"], FP32_RTOL, "sasrec dU")
    f = golden("bert4rec_ce")
    Ud, Wd, bd = (dev(T(f[k])).requires_grad_(True) for k in ("U", "W", "bias"))
    loss = ops.fused_ce(Ud, Wd, dev(T(f["labels"])), bias=bd, precision="fp32")
    loss.backward()
    assert abs(float(loss.detach()) - float(f["loss"])) <= FP32_RTOL * abs(float(f["loss"]))
    assert_rel(Ud.grad, f["dU"], FP32_RTOL, "bert4rec dU")
    assert_rel(Wd.grad, f["dW"], FP32_RTOL, "bert4rec dW")
    assert_rel(bd.grad, f["dbias"], FP32_RTOL, "bert4rec dbias")


@pytest.mark.parametrize("M,N,d,with_bias", [(300, 1000, 64, True), (513, 40_000, 128, False), (2000, 70_000, 128, True),
                                             (17000, 50_000, 64, True), (700, 30_000, 256, True), (513, 4099, 136, False)])
def test_ce_dw_bf16_output_is_the_rounded_fp32_gradient(ops, M, N, d, with_bias):
    """rb_ce_bwd_dw_bf16 (one split: direct bf16 rows + fp32 side table for label rows; several splits: fp32
    staging) must equal the fp32 gradient of rb_ce_bwd rounded to bf16, bit for bit; dbias to the last ulp."""
    g = torch.Generator().manual_seed(M + N)
    U = dev((torch.randn(M, d, generator=g) * 1.5 / d ** 0.25).bfloat16())
    W = dev((torch.randn(N, d, generator=g) * 1.5 / d ** 0.25).bfloat16())
    b = dev(torch.randn(N, generator=g) * 0.3) if with_bias else None
    lab = torch.randint(0, N, (M,), generator=g)
    lab[: M // 8] = 7        # hot label
    lab = dev(lab)
    m, l, ll = ops.ce_rowstats(U, W, lab, bias=b)
    lse = m + torch.log(l)
    _, dW32, db32 = ops.ce_backward(U, W, lab, lse, 1.0 / M, bias=b, need_dU=False, need_dW=True, need_dbias=with_bias)
    _, dWb, dbb = ops.ce_backward(U, W, lab, lse, 1.0 / M, bias=b, need_dU=False, need_dW=True, need_dbias=with_bias,
                                  dw_dtype=torch.bfloat16)
    assert dWb.dtype == torch.bfloat16
    assert torch.equal(dWb, dW32.bfloat16())
    if with_bias:   # the count term is folded in by one FMA in one path and by multiply + add in the other: 1 ulp
        assert torch.allclose(dbb, db32, rtol=1e-6, atol=1e-9)


def _seen_lists(g, B, N, max_len):
    return [torch.randperm(N, generator=g)[: int(torch.randint(0, max_len + 1, (1,), generator=g))].tolist() for _ in range(B)]


def _check_topk(vals, ids, ref_scores_masked, K, tol):
    """ids identical wherever the reference gaps exceed `tol`; values within tol everywhere."""
    rv, ri = orc.topk_sorted(ref_scores_masked, K)
    vals, ids = vals.cpu(), ids.cpu().long()
    valid = rv > orc.MASK_VALUE / 2  # entries past the unmasked catalog: the build reports (-1e23, -1)
    assert torch.all(ids[~valid] == -1) and torch.all(vals[~valid] == orc.MASK_VALUE)
    scale = float(rv[valid].abs().max()) if valid.any() else 1.0
    assert float((vals - rv)[valid].abs().max() if valid.any() else 0.0) <= tol * scale
    # a rank is "decided" when both neighbours (in the full reference ranking) are > tol away
    full_sorted = torch.sort(ref_scores_masked, dim=1, descending=True, stable=True).values
    nxt = full_sorted[:, 1 : K + 1] if full_sorted.shape[1] > K else torch.cat([full_sorted[:, 1:], full_sorted[:, -1:] - 1], 1)
    kk = rv.shape[1]
    gap_next = (rv - nxt[:, :kk]).abs() > 4 * tol * scale
    gap_prev = torch.ones_like(gap_next)
    gap_prev[:, 1:] = (rv[:, :-1] - rv[:, 1:]).abs() > 4 * tol * scale
    decided = gap_next & gap_prev & valid
    assert torch.equal(ids[decided], ri[decided])
    return float((ids == ri)[valid].float().mean()) if valid.any() else 1.0


@pytest.mark.parametrize("B,N,d,K,max_seen", [
    (1, 200, 64, 20, 10), (130, 1000, 64, 50, 40), (300, 5000, 128, 100, 64), (64, 300, 32, 100, 290),
    (257, 12101, 64, 50, 30),
])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_masked_topk_and_metrics(ops, B, N, d, K, max_seen, precision):
    if precision == "fp32" and d > 128:
        pytest.skip("fp32x3 mode supports d <= 128")
    g = torch.Generator().manual_seed(B + N)
    U = torch.randn(B, d, generator=g) / d ** 0.25
    W = torch.randn(N, d, generator=g) / d ** 0.25
    if precision == "bf16":
        U, W = bf16_round(U), bf16_round(W)
    seen = _seen_lists(g, B, N, max_seen)
    seen[0] = []  # empty seen list
    crow, col = orc.lists_to_csr(seen)
    tgt = torch.randint(0, N, (B,), generator=g)
    if seen[-1]:
        tgt[-1] = seen[-1][0]  # target inside the seen list: guaranteed miss
    tcrow, tcol = orc.lists_to_csr([[int(x)] for x in tgt])
    cast = (lambda x: dev(x).bfloat16()) if precision == "bf16" else dev
    vals, ids = ops.topk_eval(cast(U), cast(W), K, dev(crow), dev(col), precision=precision)
    masked = orc.mask_seen(orc.score_dense(U, W), crow, col)
    tol = FP32_RTOL if precision == "fp32" else 2e-6
    match = _check_topk(vals, ids, masked, K, tol)
    assert match > 0.99
    mons = [f"{m}@{k}" for m in ("HITRATE", "NDCG") for k in (1, 5, 10, 20) if k <= K]
    ref = orc.evaluate_batch(orc.score_dense(U, W), crow, col, tcrow, tcol, mons)
    got = MX.batch_metrics(ids, dev(tcrow), dev(tcol), N, mons, exact=True)
    _, ri = orc.topk_sorted(masked, K)
    if torch.equal(ids.cpu().long()[:, :20], ri[:, :20]):
        assert got == ref  # bit-identical metrics whenever the ranked ids agree
    else:
        assert all(abs(got[k] - ref[k]) <= 2.0 / B for k in ref)


@pytest.mark.parametrize("B,N,d,K,offset", [(300, 5000, 64, 20, 0), (513, 40_000, 128, 50, 1), (200, 3000, 256, 10, 0)])
def test_masked_topk_with_bias_head(ops, B, N, d, K, offset):
    """BERT4Rec-style head (scale * U W^T + bias, BERT4Rec/main.py:83,189): masked top-K and row statistics with a bias
    vector that is 16-byte aligned (vector loads in the epilogue) or not (``offset`` = 1: a slice of a larger buffer)."""
    g = torch.Generator().manual_seed(B + N + d)
    U = bf16_round(torch.randn(B, d, generator=g) / d ** 0.25)
    W = bf16_round(torch.randn(N, d, generator=g) / d ** 0.25)
    bias = torch.randn(N, generator=g) * 0.5
    lab = torch.randint(0, N, (B,), generator=g)
    seen = _seen_lists(g, B, N, 40)
    crow, col = orc.lists_to_csr(seen)
    bd = dev(torch.cat([torch.zeros(offset), bias]))[offset:]
    vals, ids = ops.topk_eval(dev(U).bfloat16(), dev(W).bfloat16(), K, dev(crow), dev(col), bias=bd, scale=0.7)
    masked = orc.mask_seen(orc.score_dense(U, W, bias, 0.7), crow, col)
    assert _check_topk(vals, ids, masked, K, 2e-6) > 0.99
    m, l, ll = ops.ce_rowstats(dev(U).bfloat16(), dev(W).bfloat16(), dev(lab), bias=bd, scale=0.7)
    rm, rl, rll = orc.ce_rowstats(U, W, lab, bias, 0.7)
    lse, rlse = (m + torch.log(l)).cpu(), rm + torch.log(rl)
    assert float((lse - rlse).abs().max()) <= 1e-5 * float(rlse.abs().max()) + 1e-6
    assert_rel(ll, rll, 1e-5, "label logit")


def test_golden_unisrec_evaluate_masked_topk_and_hits(ops, golden):
    """a8-a10 against the reference's own evaluate (tests/golden/unisrec_evaluate.npz: the masked scores and dense
    targets ``CoachForUniSRec.evaluate`` handed to its metric functions): the fused masked top-K must be the top-K of
    those masked scores, the hit matrix must be ``targets.gather(1, topk ids)``."""
    g = golden("unisrec_evaluate")
    W = T(g["item_table"])
    K = 20
    for b in range(int(g["n_batches"])):
        U = T(g[f"U{b}"])
        masked, targets = T(g[f"scores_masked{b}"]), T(g[f"targets{b}"])
        vals, ids = ops.topk_eval(dev(U), dev(W), K, dev(T(g[f"seen_crow{b}"])), dev(T(g[f"seen_col{b}"])), precision="fp32")
        rv, ri = orc.topk_sorted(masked, K)
        assert_rel(vals, rv, FP32_RTOL, "masked top-K values")
        assert torch.equal(ids.cpu().long(), ri)
        hits = MX.hits_from_topk(ids, dev(T(g[f"tgt_crow{b}"])), dev(T(g[f"tgt_col{b}"])), W.shape[0])
        assert torch.equal(hits.cpu(), targets.gather(1, ri))


@pytest.mark.parametrize("B,N,K,multi", [(300, 5000, 50, False), (257, 900, 100, True), (1, 40, 7, True)])
def test_metrics_in_one_pass_match_the_exact_reduction(ops, B, N, K, multi):
    """rb_topk_metrics (all METRIC@k of a batch from the ranked ids, on the device) against the oracle's float32
    reductions: single-target rows (LOU) and ragged multi-target rows incl. rows without targets and -1 ids."""
    g = torch.Generator().manual_seed(B + K)
    ids = torch.stack([torch.randperm(N, generator=g)[:K] for _ in range(B)]).int()
    ids[0, K // 2:] = -1                         # a row whose ranked list ran out of unmasked items
    tgts = []
    for r in range(B):
        n_t = int(torch.randint(0, 6, (1,), generator=g)) if multi else 1
        pick = ids[r, torch.randperm(K, generator=g)[: n_t // 2 + 1]].tolist() if r % 3 else []
        extra = torch.randint(0, N, (max(n_t - len(pick), 0),), generator=g).tolist()
        t = sorted(set(int(x) for x in pick + extra if x >= 0))
        tgts.append(t if (multi or t) else [int(torch.randint(0, N, (1,), generator=g))])
    crow, col = orc.lists_to_csr(tgts)
    mons = [f"{m}@{k}" for m in ("HITRATE", "RECALL", "PRECISION", "NDCG", "MRR") for k in (1, 5, K) if k <= K]
    n_t = (crow[1:] - crow[:-1]).float()
    ref = MX.metrics_from_hits(orc.hits_from_topk(ids, crow, col, N), n_t, mons)
    got = MX.batch_metrics(dev(ids), dev(crow), dev(col), N, mons, exact=False)
    assert set(got) == set(ref)
    for k in ref:
        assert abs(got[k] - ref[k]) <= 1e-6 * max(1.0, abs(ref[k])), (k, got[k], ref[k])


def test_topk_hits_kernel_matches_host_logic(ops):
    """rb_topk_hits vs the oracle's dense-target gather (multi-target rows, missing entries, empty rows)."""
    g = torch.Generator().manual_seed(3)
    B, K, N = 300, 50, 1000
    ids = torch.randint(0, N, (B, K), generator=g).int()
    ids[5, 40:] = -1
    tl = [torch.randperm(N, generator=g)[: int(torch.randint(0, 6, (1,), generator=g))].tolist() for _ in range(B)]
    for b in range(0, B, 3):
        if tl[b]:
            ids[b, b % K] = tl[b][0]      # plant hits
    crow, col = orc.lists_to_csr(tl)
    ref = orc.hits_from_topk(ids, crow, col, N)
    got = MX.hits_from_topk(dev(ids), dev(crow), dev(col), N)
    assert torch.equal(got.cpu(), ref) and float(ref.sum()) > 0


def test_topk_without_seen_and_ties(ops):
    # duplicated item rows => exact score ties: lowest id must win, like the oracle's stable sort
    g = torch.Generator().manual_seed(9)
    U = bf16_round(torch.randn(40, 64, generator=g))
    Wb = bf16_round(torch.randn(100, 64, generator=g))
    W = torch.cat([Wb, Wb, Wb])  # ids i, i+100, i+200 tie
    vals, ids = ops.topk_eval(dev(U).bfloat16(), dev(W).bfloat16(), 30)
    rv, ri = orc.topk_sorted(orc.score_dense(U, W), 30)
    assert torch.equal(ids.cpu().long(), ri)
    assert_rel(vals, rv, 2e-6, "tie vals")


def test_topk_heavy_users_keep_a_threshold(ops):
    """Users whose seen list covers a large part of a small catalog (every 128-item tile holds seen ids): the prefix
    maxima run over the unseen items and candidate groups are dirty one by one, so those rows still get a threshold and
    do not fall back to the exact scan (checked through rb_topk_debug_layout: no overflowed sub-list)."""
    import ctypes
    from recboard_b200 import _lib as L
    g = torch.Generator().manual_seed(33)
    B, N, d, K = 300, 38_048, 64, 20
    U = bf16_round(torch.randn(B, d, generator=g) / d ** 0.25)
    W = bf16_round(torch.randn(N, d, generator=g) / d ** 0.25)
    seen = _seen_lists(g, B, N, 40)
    for r in range(0, B, 7):                       # heavy users: 3000-9000 seen ids, ~10-30 per tile
        seen[r] = torch.randperm(N, generator=g)[: 3000 + 20 * r].tolist()
    crow, col = orc.lists_to_csr(seen)
    vals, ids = ops.topk_eval(dev(U).bfloat16(), dev(W).bfloat16(), K, dev(crow), dev(col))
    masked = orc.mask_seen(orc.score_dense(U, W), crow, col)
    assert _check_topk(vals, ids, masked, K, 2e-6) > 0.99
    out = (ctypes.c_int64 * 8)()
    L.check(L.lib().rb_topk_debug_layout(B, N, d, K, L.MODE_BF16, col.numel(), out), "rb_topk_debug_layout")
    ws = next(iter(L.Workspace._bufs.values()))
    overflow = ws[out[5]:out[5] + 4 * B].view(torch.int32)
    assert int(overflow.sum()) == 0, f"{int(overflow.sum())} rows fell back to the exact scan"


def test_topk_candidate_overflow_falls_back(ops):
    # massive ties: every item of a 5000-item catalog scores the same for half of the rows, so the candidate
    # list of the threshold sweep overflows and the exact fallback must take over (lowest ids win)
    g = torch.Generator().manual_seed(10)
    B, N, d, K = 70, 5000, 64, 20
    U = bf16_round(torch.randn(B, d, generator=g))
    W = bf16_round(torch.randn(N, d, generator=g))
    U[::2] = 0.0                                  # all scores == 0 for these rows
    W[1000:4000] = W[999]                         # and a 3000-item plateau for the others
    seen = [[0, 1, 2, 1500] for _ in range(B)]
    crow, col = orc.lists_to_csr(seen)
    vals, ids = ops.topk_eval(dev(U).bfloat16(), dev(W).bfloat16(), K, dev(crow), dev(col))
    masked = orc.mask_seen(orc.score_dense(U, W), crow, col)
    rv, ri = orc.topk_sorted(masked, K)
    assert torch.equal(ids.cpu().long()[::2], ri[::2])            # exact ties: ids 3,4,5,...
    assert torch.equal(ids.cpu().long()[0], torch.arange(3, 3 + K))
    assert_rel(vals, rv, 2e-6, "overflow vals")
    _check_topk(vals, ids, masked, K, 2e-6)


@pytest.mark.parametrize("M,K,N,d,dt", [(3, 1, 50, 64, torch.float32), (300, 101, 4000, 64, torch.float32),
                                         (513, 513, 12101, 128, torch.bfloat16), (64, 7, 900, 96, torch.bfloat16),
                                         (40, 33, 700, 256, torch.float32)])
def test_gather_dot_forward_backward(ops, M, K, N, d, dt):
    """pool / sampled-softmax scoring without the (M,K,d) gather; hot ids, out-of-range ids and
    a padding row included.  fp32 storage 1e-5, bf16 storage: same inputs as the oracle, fp32 accumulate."""
    g = torch.Generator().manual_seed(M * 31 + K)
    U = (torch.randn(M, d, generator=g) / d ** 0.25).to(dt)
    tab = (torch.randn(N, d, generator=g) / d ** 0.25).to(dt)
    idx = torch.randint(0, N, (M, K), generator=g)
    idx[:, 0] = 5                       # hot row: every query hits it
    if K > 2:
        idx[0, 1] = 0                   # the padding row
    Gup = torch.randn(M, K, generator=g)
    Ur, Tr = U.float().clone().requires_grad_(True), tab.float().clone().requires_grad_(True)
    ref = orc.gather_dot(Ur, Tr, idx, 0.7)
    (ref * Gup).sum().backward()
    rdT = Tr.grad.clone()
    rdT[0] = 0                          # padding_idx = 0 keeps row 0 at zero (nn.Embedding contract)
    Ud, Td = dev(U).requires_grad_(True), dev(tab).requires_grad_(True)
    S = ops.gather_dot(Ud, Td, dev(idx), scale=0.7, padding_idx=0)
    (S * dev(Gup)).sum().backward()
    tol = FP32_RTOL
    assert_rel(S, ref.detach(), tol, "scores")
    assert_rel(Ud.grad, Ur.grad, tol if dt == torch.float32 else 2.0 ** -8, "dU")
    assert_rel(Td.grad, rdT, tol if dt == torch.float32 else 2.0 ** -8, "dTable")
    # determinism: the table gradient is bitwise reproducible
    Ud2, Td2 = dev(U).requires_grad_(True), dev(tab).requires_grad_(True)
    (ops.gather_dot(Ud2, Td2, dev(idx), scale=0.7, padding_idx=0) * dev(Gup)).sum().backward()
    assert torch.equal(Td.grad, Td2.grad) and torch.equal(Ud.grad, Ud2.grad)


@pytest.mark.parametrize("R,d,avg_deg", [(1, 64, 3), (700, 64, 9), (2049, 128, 40), (300, 32, 0), (513, 8, 5)])
def test_spmm_csr(ops, R, d, avg_deg):
    """LightGCN propagation (LightGCN/main.py:83-85), forward, fused average and backward."""
    g = torch.Generator().manual_seed(R + d)
    nnz = R * avg_deg
    r = torch.randint(0, R, (nnz,), generator=g)
    c = torch.randint(0, R, (nnz,), generator=g)
    v = torch.rand(nnz, generator=g)
    A = torch.sparse_coo_tensor(torch.stack([torch.cat([r, c]), torch.cat([c, r])]), torch.cat([v, v]), (R, R)).coalesce().to_sparse_csr()
    X = torch.randn(R, d, generator=g)
    Xr = X.clone().requires_grad_(True)
    ref = orc.spmm(A, Xr)
    Gup = torch.randn(R, d, generator=g)
    (ref * Gup).sum().backward()
    Xd = dev(X).requires_grad_(True)
    Y = ops.spmm(dev(A), Xd, symmetric=True)
    (Y * dev(Gup)).sum().backward()
    assert_rel(Y, ref.detach(), FP32_RTOL, "A @ X")
    assert_rel(Xd.grad, Xr.grad, FP32_RTOL, "dX")
    acc = dev(X).clone()
    ops.spmm_raw(dev(A), dev(X), acc=acc, beta=0.25, want_y=False)
    assert_rel(acc, X + 0.25 * ref.detach(), FP32_RTOL, "fused average")
    # general (non-symmetric) matrices take the transposed CSR in backward
    B = torch.sparse_coo_tensor(torch.stack([r, c]), v, (R, R)).coalesce().to_sparse_csr() if nnz else A
    Xr2 = X.clone().requires_grad_(True)
    (orc.spmm(B, Xr2) * Gup).sum().backward()
    Xd2 = dev(X).requires_grad_(True)
    (ops.spmm(dev(B), Xd2) * dev(Gup)).sum().backward()
    assert_rel(Xd2.grad, Xr2.grad, FP32_RTOL, "dX (general)")


@pytest.mark.parametrize("n,d,in_dt,out_dt", [(1, 8, torch.float32, torch.float32), (1000, 256, torch.float32, torch.bfloat16),
                                             (4097, 64, torch.bfloat16, torch.bfloat16), (333, 1024, torch.bfloat16, torch.float32)])
def test_normalize_rows(ops, n, d, in_dt, out_dt):
    """a11: F.normalize(weight[1:], dim=-1) (HSTU/main.py:182-184), including an all-zero row (eps clamp)."""
    g = torch.Generator().manual_seed(n + d)
    x = (torch.randn(n + 1, d, generator=g) * 3).to(in_dt)
    x[n // 2 + 1] = 0
    out, inv = ops.normalize_rows(dev(x)[1:], out_dtype=out_dt, return_inv_norm=True)   # the weight[NUM_PADS:] view
    ref = orc.normalize_rows(x[1:].float())
    assert out.dtype == out_dt and out.shape == (n, d)
    tol = 1e-6 if out_dt == torch.float32 else 2.0 ** -8
    assert float((out.float().cpu() - ref).abs().max()) <= tol
    rn = x[1:].float().norm(dim=-1).clamp_min(1e-12)
    nz = rn > 1e-6
    assert torch.allclose(inv.cpu()[nz], 1.0 / rn[nz], rtol=1e-5)
    assert torch.all(out[n // 2] == 0)


def test_scatter_add_hot_rows_deterministic(ops):
    g = torch.Generator().manual_seed(3)
    n_rows, d, n = 5000, 128, 40000
    idx = (torch.rand(n, generator=g) ** 6 * n_rows).long()  # heavy duplicates (popularity skew)
    idx[:100] = 0
    go = torch.randn(n, d, generator=g)
    ref = orc.scatter_add_rows(go, idx, n_rows, padding_idx=0)
    outs = []
    for _ in range(2):
        t = torch.zeros(n_rows, d, device="cuda")
        ops.scatter_add_rows_(t, dev(go), dev(idx), padding_idx=0)
        outs.append(t)
    assert torch.equal(outs[0], outs[1])
    assert_rel(outs[0], ref, 1e-5, "scatter-add")
    tb = torch.zeros(n_rows, d, device="cuda")
    ops.scatter_add_rows_(tb, dev(go).bfloat16(), dev(idx), padding_idx=0)
    assert_rel(tb, orc.scatter_add_rows(bf16_round(go), idx, n_rows, 0), 1e-5, "scatter-add bf16")


@pytest.mark.parametrize("n_rows,n", [(5000, 40000), (1_000_001, 204_800), (5_000_000, 7), (300, 1)])
def test_scatter_add_sizes_and_bf16_table(ops, n_rows, n):
    """The one-launch scatter-add at the bench size (two 11-bit passes), beyond 4M rows (three passes), tiny inputs;
    into an fp32 table that already holds values, and into a bf16 table (fp32 add, one rounding per touched row)."""
    g = torch.Generator().manual_seed(n)
    d = 64
    idx = (torch.rand(n, generator=g) ** 4 * n_rows).long().clamp_(0, n_rows - 1)
    go = torch.randn(n, d, generator=g).bfloat16()
    add = orc.scatter_add_rows(go.float(), idx, n_rows, padding_idx=0)
    touched = torch.unique(idx)
    base = torch.zeros(n_rows, d)
    base[touched] = torch.randn(len(touched), d, generator=g)
    t32 = dev(base.clone())
    ops.scatter_add_rows_(t32, dev(go), dev(idx), padding_idx=0)
    got = t32[dev(touched)].cpu()
    assert_rel(got, (base + add)[touched], 1e-5, "accumulate into fp32")
    untouched = torch.ones(n_rows, dtype=torch.bool); untouched[touched] = False
    assert bool((t32.cpu()[untouched] == 0).all())
    tb = dev(base.bfloat16())
    ops.scatter_add_rows_(tb, dev(go), dev(idx), padding_idx=0)
    ref_b = (base.bfloat16().float() + add).bfloat16()
    assert torch.equal(tb[dev(touched)].cpu(), ref_b[touched]) or \
        float((tb[dev(touched)].cpu().float() - ref_b[touched].float()).abs().max()) <= 2 ** -7 * float(ref_b[touched].float().abs().max())


def test_gather_rows_backward_accumulates_into_existing_grad(ops):
    """autograd of the gather (SASRec/main.py:183 at :249): without a gradient buffer a dense gradient in the
    parameter's dtype is returned; with ``accumulate=True`` and an existing ``.grad`` the rows are added in place --
    and the sum with the scoring head's dW (fused_ce with n_skip) is the reference's single (N+P,d) gradient."""
    g = torch.Generator().manual_seed(4)
    N, P, d, B, S, M = 900, 1, 64, 16, 10, 50
    W0 = torch.randn(N + P, d, generator=g) * 0.3
    W0[0] = 0
    idx = torch.randint(0, N + P, (B, S), generator=g)
    U = torch.randn(M, d, generator=g) * 0.3
    labels = torch.randint(0, N, (M,), generator=g)
    # reference: plain torch autograd on the CPU (nn.Embedding with padding_idx + einsum + cross_entropy)
    Wr = W0.clone().requires_grad_(True)
    emb = torch.nn.functional.embedding(idx, Wr, padding_idx=0)
    loss = torch.nn.functional.cross_entropy(U @ Wr[P:].T, labels) + (emb * emb).sum() * 0.01
    loss.backward()
    for accumulate in (False, True):
        Wd = dev(W0).requires_grad_(True)
        if accumulate:
            Wd.grad = torch.zeros_like(Wd)
        buf = Wd.grad
        embd = ops.gather_rows(Wd, dev(idx), padding_idx=0, accumulate=accumulate)
        lossd = ops.fused_ce(dev(U), Wd, dev(labels), n_skip=P) + (embd * embd).sum() * 0.01
        lossd.backward()
        assert abs(float(lossd) - float(loss)) <= 1e-5 * abs(float(loss))
        assert_rel(Wd.grad, Wr.grad, 2e-5, f"table gradient (accumulate={accumulate})")
        assert bool((Wd.grad[0] == 0).all())                     # the padding row stays zero
        if accumulate:
            assert Wd.grad.data_ptr() == buf.data_ptr()           # still the same buffer: nothing was re-allocated


@pytest.mark.parametrize("d", [128, 256])
def test_bf16_table_gradient_built_in_one_buffer(ops, d):
    """bf16 parameter, ``.grad`` kept allocated: the gather's rows and the head's dW are both added inside their own
    kernels into that ONE (N+P,d) buffer (gather_rows(accumulate=True) + fused_ce(n_skip, accumulate=True))."""
    g = torch.Generator().manual_seed(6)
    N, P, B, S, M = 3000, 1, 32, 12, 300
    W0 = bf16_round(torch.randn(N + P, d, generator=g) * 0.3)
    W0[0] = 0
    idx = torch.randint(0, N + P, (B, S), generator=g)
    U = bf16_round(torch.randn(M, d, generator=g) * 0.3)
    labels = torch.randint(0, N, (M,), generator=g)
    gemb = bf16_round(torch.randn(B, S, d, generator=g) * 0.01)
    Wr = W0.clone().requires_grad_(True)
    emb = torch.nn.functional.embedding(idx, Wr, padding_idx=0)
    loss = torch.nn.functional.cross_entropy(U @ Wr[P:].T, labels)
    torch.autograd.backward([loss, emb], [None, gemb])
    old = bf16_round(torch.randn(N + P, d, generator=g) * 1e-3)
    Wd = dev(W0).bfloat16().requires_grad_(True)
    Wd.grad = dev(old).bfloat16()
    buf = Wd.grad
    Ud = dev(U).bfloat16().requires_grad_(True)
    embd = ops.gather_rows(Wd, dev(idx), padding_idx=0, accumulate=True)
    lossd = ops.fused_ce(Ud, Wd, dev(labels), n_skip=P, accumulate=True)
    torch.autograd.backward([lossd, embd], [None, dev(gemb).bfloat16()])
    assert Wd.grad.data_ptr() == buf.data_ptr()
    assert abs(float(lossd) - float(loss)) <= 1e-5 * abs(float(loss))
    assert_rel(Wd.grad.float() - dev(old), Wr.grad, 3 * BF16_RTOL, "bf16 table gradient (two roundings: gather rows, dW)")
    assert bool((Wd.grad[0].float().cpu() == old[0]).all())     # the padding row received nothing
    # after zero_grad() (set_to_none=True): the dW pass writes the new gradient buffer, the gather adds into it
    Wd.grad = None
    Ud2 = dev(U).bfloat16().requires_grad_(True)
    embd = ops.gather_rows(Wd, dev(idx), padding_idx=0, accumulate=True)
    lossd = ops.fused_ce(Ud2, Wd, dev(labels), n_skip=P, accumulate=True)
    torch.autograd.backward([lossd, embd], [None, dev(gemb).bfloat16()])
    assert Wd.grad is not None and Wd.grad.dtype == torch.bfloat16 and Wd.grad.shape == Wd.shape
    assert_rel(Wd.grad.float(), Wr.grad, 3 * BF16_RTOL, "bf16 table gradient from a fresh buffer")
    assert float(Wd.grad[0].float().abs().max()) == 0.0
    assert_grad_bf16(Ud2.grad, Ud.grad.float().cpu(), "dU unchanged")


def test_config2_full_size_topk_against_fp64_oracle(ops):
    """BASELINE configs[1] at FULL size -- all 31 668 users x 38 048 items, d = 64, fp32 parity mode, top-20 with seen
    masking in one call -- against the oracle lines (score_dense + mask_seen + topk, UniSRec/main.py:408-435) evaluated
    in float64 on the GPU, 4096 users at a time."""
    from recboard_b200 import synth
    B, N, d, K = 31668, 38048, 64, 20
    g = torch.Generator(device="cuda").manual_seed(2028)
    U = synth.embeddings(B, d, g, torch.device("cuda"), torch.float32, gain=1.5)
    W = synth.embeddings(N, d, g, torch.device("cuda"), torch.float32, gain=1.5)
    crow, col = synth.seen_csr(B, N, g, torch.device("cuda"))
    vals, ids = ops.topk_eval(U, W, K, crow, col, precision="fp32")
    bad_vals = bad_ids = 0
    for lo in range(0, B, 4096):
        hi = min(lo + 4096, B)
        a, b = int(crow[lo]), int(crow[hi])
        S = orc.mask_seen(orc.score_dense(U[lo:hi].double(), W.double()), crow[lo:hi + 1] - a, col[a:b])
        rv, ri = torch.topk(S, K + 1, dim=1)
        scale = float(rv[:, :K].abs().max())
        bad_vals += int(((vals[lo:hi].double() - rv[:, :K]).abs() > FP32_RTOL * scale).sum())
        gap = (rv[:, :-1] - rv[:, 1:]) > 4 * FP32_RTOL * scale           # neighbours fp32 arithmetic can tell apart
        gap_prev = torch.cat([torch.ones_like(gap[:, :1]), gap[:, :K - 1]], 1)
        bad_ids += int(((ids[lo:hi].long() != ri[:, :K]) & gap[:, :K] & gap_prev).sum())
    assert bad_vals == 0 and bad_ids == 0, (bad_vals, bad_ids)


def test_config3_full_size_gradients_against_fp64_oracle(ops):
    """BASELINE configs[2] at FULL size (4096 x 1M x 128, bf16 operands): loss, the whole dU and the whole dW of the
    fused passes against the oracle's closed-form CE gradients in float64 (chunked over query rows, on the GPU)."""
    from recboard_b200 import synth
    M, N, d = 4096, 1_000_000, 128
    cu = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(2029)
    U = synth.embeddings(M, d, g, cu, torch.bfloat16, gain=1.5)
    W = synth.embeddings(N, d, g, cu, torch.bfloat16, gain=1.5)
    labels = synth.zipf_ids(M, N, g, cu)
    ref_loss, rdU, rdW, _ = orc.ce_fwd_bwd_chunked(U.double(), W.double(), labels, chunk=256)
    Ud, Wd = U.clone().requires_grad_(True), W.clone().requires_grad_(True)
    loss = ops.fused_ce(Ud, Wd, labels)
    loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    for got, ref, name in ((Ud.grad, rdU, "dU"), (Wd.grad, rdW, "dW")):
        mx = float(ref.abs().max())
        err = float((got.double() - ref).abs().max()) / mx
        assert err <= BF16_RTOL + 2.0 ** -9, (name, err)      # north_star tolerance + the bf16 storage of the gradient
    # fp32 gradient outputs (no storage rounding): the bare 2e-3
    m_, l_, ll_, du = ops.ce_rowstats(U, W, labels, want_dU=True)
    lse = m_ + torch.log(l_)
    dU32 = ops.ce_du_finish(du, m_, lse, W, labels, 1.0 / M)
    _, dW32, _ = ops.ce_backward(U, W, labels, lse, 1.0 / M, need_dU=False, need_dW=True)
    for got, ref, name in ((dU32, rdU, "dU fp32"), (dW32, rdW, "dW fp32")):
        err = float((got.double() - ref).abs().max()) / float(ref.abs().max())
        assert err <= BF16_RTOL, (name, err)


@pytest.mark.parametrize("precision,d", [("fp32", 64), ("bf16", 64), ("bf16", 256)])
def test_device_side_row_count_matches_compacted_rows(ops, precision, d):
    """a4: capacity rows + a device-side count (ops.compact_queries -> fused_ce(n_valid=...)) against the oracle on the
    rows torch's boolean indexing selects -- loss, dX through the compaction, dW -- with NO host synchronisation
    between the mask and the loss (torch.cuda.set_sync_debug_mode("error"))."""
    g = torch.Generator().manual_seed(17)
    B, S, N = 24, 30, 2000
    X = torch.randn(B, S, d, generator=g) * 0.4 * (64 / d) ** 0.25
    W = torch.randn(N, d, generator=g) * 0.4 * (64 / d) ** 0.25
    pos = torch.randint(0, N, (B, S), generator=g)
    mask = torch.rand(B, S, generator=g) < 0.2
    mask[0, 0] = True
    if precision == "bf16":
        X, W = bf16_round(X), bf16_round(W)
    Xr = X.clone().requires_grad_(True)
    ref_loss, _, ref_dW, _ = orc.ce_fwd_bwd(X[mask], W, pos[mask])
    lossr = orc.ce_loss(Xr[mask], W, pos[mask])
    lossr.backward()
    cast = (lambda x: dev(x).bfloat16()) if precision == "bf16" else dev
    Xd, Wd = cast(X).requires_grad_(True), cast(W).requires_grad_(True)
    maskd, posd = dev(mask), dev(pos)
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        U, (labels,), count = ops.compact_queries(Xd, maskd, posd)
        loss = ops.fused_ce(U, Wd, labels, n_valid=count, precision=precision)
        loss.backward()
    finally:
        torch.cuda.set_sync_debug_mode("default")
    assert int(count) == int(mask.sum()) and U.shape[0] == B * S
    tol = FP32_RTOL if precision == "fp32" else 1e-5
    assert abs(float(loss) - float(ref_loss)) <= tol * abs(float(ref_loss))
    if precision == "fp32":
        assert_rel(Xd.grad, Xr.grad, 2e-5, "dX through the compaction")
        assert_rel(Wd.grad, ref_dW, 2e-5, "dW")
    else:
        assert_grad_bf16(Xd.grad, Xr.grad, "dX through the compaction")
        assert_grad_bf16(Wd.grad, ref_dW, "dW")
    # nothing selected at all: F.cross_entropy(mean) over zero rows is NaN, the gradients are zero
    Xz, Wz = cast(X).requires_grad_(True), cast(W).requires_grad_(True)
    U0, (lab0,), c0 = ops.compact_queries(Xz, torch.zeros_like(maskd), posd)
    l0 = ops.fused_ce(U0, Wz, lab0, n_valid=c0, precision=precision, reduction="sum")
    l0.backward()
    assert float(l0) == 0.0 and int(c0) == 0
    assert float(Wz.grad.float().abs().max()) == 0.0 and float(Xz.grad.float().abs().max()) == 0.0


def test_gather_over_several_shards_one_device(ops):
    """rb_gather_rows_peers with the shards of a table as separate allocations (here all on this GPU; across ranks they
    are IPC-mapped peers, tests/multi_gpu_check.py): every id is served by the shard that owns it, ids outside the
    table give zero rows; rb_ipc_export returns the allocation handle and the offset of an interior pointer."""
    import ctypes as C
    from recboard_b200 import _lib as L
    g = torch.Generator().manual_seed(12)
    N, d = 10_007, 64
    W = bf16_round(torch.randn(N, d, generator=g))
    bounds = [0, 3000, 3001, 7777, N]                       # ragged shards, one of a single row
    shards = [dev(W[a:b]).bfloat16().contiguous() for a, b in zip(bounds[:-1], bounds[1:])]
    idx = torch.randint(-3, N + 5, (500, 7), generator=g)   # incl. ids below and above the table
    idx[0, :4] = torch.tensor([0, 2999, 3000, N - 1])
    out = torch.empty(idx.numel(), d, dtype=torch.bfloat16, device="cuda")
    ptrs = (C.c_void_p * len(shards))(*[s_.data_ptr() for s_ in shards])
    starts = (C.c_int64 * len(bounds))(*bounds)
    L.call(out.device, "rb_gather_rows_peers", ptrs, starts, len(shards), L.ptr(dev(idx).view(-1)), L.ptr(out), idx.numel(), d,
           L.DTYPE_BF16, L.stream_ptr(out.device))
    ok = (idx >= 0) & (idx < N)
    ref = torch.where(ok.unsqueeze(-1), W[idx.clamp(0, N - 1)], torch.zeros(1, d)).view(-1, d)
    assert torch.equal(out.float().cpu(), ref)
    handle, off = C.create_string_buffer(64), C.c_int64(-1)
    inner = shards[3][5:]                                   # a pointer inside an allocation
    L.call(out.device, "rb_ipc_export", L.ptr(inner), C.cast(handle, C.c_void_p), C.cast(C.pointer(off), C.c_void_p))
    assert off.value >= 5 * d * 2 and any(handle.raw)


def test_sharded_partials_merge_like_multi_gpu(ops):
    """Single-GPU simulation of R row shards: sharded stats/top-K/dW == unsharded."""
    from recboard_b200 import sharded
    g = torch.Generator().manual_seed(21)
    M, N, d, K, R = 200, 3001, 64, 20, 3
    U = bf16_round(torch.randn(M, d, generator=g) / d ** 0.25)
    W = bf16_round(torch.randn(N, d, generator=g) / d ** 0.25)
    lab = torch.randint(0, N, (M,), generator=g)
    Ud, Wd, labd = dev(U).bfloat16(), dev(W).bfloat16(), dev(lab)
    stats, tops = [], []
    for r in range(R):
        a, b = sharded.shard_bounds(N, R, r)
        stats.append(torch.stack(ops.ce_rowstats(Ud, Wd[a:b].contiguous(), labd, label_base=a)))
        tops.append(ops.topk_eval(Ud, Wd[a:b].contiguous(), K, id_base=a))
    lse, ll = sharded.merge_rowstats(torch.stack(stats))
    rm, rl, rll = orc.ce_rowstats(U, W, lab)
    assert float((lse.cpu() - (rm + torch.log(rl))).abs().max()) < 1e-5
    assert_rel(ll, rll, 1e-5, "label logit")
    mv, mi = ops.topk_merge(torch.stack([t[0] for t in tops]), torch.stack([t[1] for t in tops]))
    gv, gi = ops.topk_eval(Ud, Wd, K)
    # every finishing path re-scores its winners with the same fp32 FMA chain: bit-identical values
    assert torch.equal(mi, gi) and torch.equal(mv, gv)
    # the one-launch merges of the multi-GPU path: row statistics, and the top-K lists in the packed layout one
    # all-gather leaves behind (R,2,B,K)
    lse_k, ll_k = ops.rowstats_merge(torch.stack(stats))
    assert torch.allclose(lse_k, lse, rtol=0, atol=2e-6) and torch.allclose(ll_k, ll, rtol=1e-6, atol=1e-6)
    packed = torch.stack([torch.stack([t[0].view(torch.int32), t[1]]) for t in tops]).contiguous()
    pv, pi = ops.topk_merge_packed(packed)
    assert torch.equal(pi, mi) and torch.equal(pv, mv)
    # gradient shards: dW of shard == rows of the unsharded dW; partial dU sum == dU
    _, rdU, rdW, _ = orc.ce_fwd_bwd(U, W, lab)
    dU_sum = torch.zeros(M, d, device="cuda")
    for r in range(R):
        a, b = sharded.shard_bounds(N, R, r)
        dU, dW, _ = ops.ce_backward(Ud, Wd[a:b].contiguous(), labd, lse, 1.0 / M, label_base=a)
        dU_sum += dU
        assert_rel(dW, rdW[a:b], BF16_RTOL * float(rdW.abs().max() / rdW[a:b].abs().max()), f"dW shard {r}")
    assert_grad_bf16(dU_sum, rdU, "dU")
    # the fused forward+dU sweep: per-shard unnormalised accumulators finish against the global lse
    dU_sum2 = torch.zeros(M, d, device="cuda")
    for r in range(R):
        a, b = sharded.shard_bounds(N, R, r)
        Ws = Wd[a:b].contiguous()
        m_, l_, ll_, du_un = ops.ce_rowstats(Ud, Ws, labd, label_base=a, want_dU=True)
        assert torch.allclose((m_ + torch.log(l_)), stats[r][0] + torch.log(stats[r][1]), rtol=0, atol=2e-5)
        dU_sum2 += ops.ce_du_finish(du_un, m_, lse, Ws, labd, 1.0 / M, label_base=a)
    assert_grad_bf16(dU_sum2, rdU, "dU (fused forward)")


def test_fused_forward_lazy_reference_moves(ops):
    """Rows whose maximum keeps growing by more than the rescale threshold from tile to tile (and a
    huge first tile followed by small ones): the lazily updated reference must stay exact."""
    g = torch.Generator().manual_seed(11)
    M, N, d = 260, 128 * 9 + 17, 64
    U = bf16_round(torch.randn(M, d, generator=g) / d ** 0.5)
    W = bf16_round(torch.randn(N, d, generator=g) / d ** 0.5)
    ramp = (torch.arange(N) // 128).float() * 3.0          # item norms grow tile by tile
    W = bf16_round(W * (1.0 + ramp).unsqueeze(1))
    U = bf16_round(U * 6.0)
    W[:128] = bf16_round(W[:128] * torch.where(torch.arange(128) % 2 == 0, 40.0, 1.0).unsqueeze(1))  # huge first tile
    lab = torch.randint(0, N, (M,), generator=g)
    ref_loss, rdU, rdW, _ = orc.ce_fwd_bwd(U, W, lab)
    Ud, Wd = dev(U).bfloat16().requires_grad_(True), dev(W).bfloat16().requires_grad_(True)
    loss = ops.fused_ce(Ud, Wd, dev(lab))
    loss.backward()
    assert torch.isfinite(loss)
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    assert_grad_bf16(Ud.grad, rdU, "dU")
    assert_grad_bf16(Wd.grad, rdW, "dW")


def test_full_size_properties(ops):
    """BASELINE configs[2] size (M=4096, N=1M, d=128, bf16): size-independent properties."""
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, d, K = 4096, 1_000_000, 128, 50
    U = (torch.randn(M, d, generator=g, device="cuda") * 1.5 / d ** 0.25).bfloat16()
    W = (torch.randn(N, d, generator=g, device="cuda") * 1.5 / d ** 0.25).bfloat16()
    lab = torch.randint(0, N, (M,), generator=g, device="cuda")
    m, l, ll = ops.ce_rowstats(U, W, lab)
    lse = m + torch.log(l)
    # (1) label logit == direct dot product; (2) lse >= max and lse <= max + log N
    direct = (U.float() * W[lab].float()).sum(-1)
    assert_rel(ll, direct.cpu(), 1e-5, "label logit")
    assert torch.all(lse >= m) and torch.all(lse <= m + np.log(N) + 1e-3)
    # (3) rows of softmax - onehot sum to zero  =>  sum_j dW_j . 1 relation: sum over items of dW == sum_i (sum_j G_ij) u_i = 0
    dU, dW, _ = ops.ce_backward(U, W, lab, lse, 1.0 / M)
    col_sum = dW.double().sum(0)
    ref_scale = float(dW.double().abs().sum(0).max())
    assert float(col_sum.abs().max()) <= 5e-3 * ref_scale
    # (4) split-invariance: stats of two half catalogs merge to the full stats
    from recboard_b200 import sharded
    h = N // 2
    s0 = torch.stack(ops.ce_rowstats(U, W[:h], lab, label_base=0))
    s1 = torch.stack(ops.ce_rowstats(U, W[h:], lab, label_base=h))
    lse2, ll2 = sharded.merge_rowstats(torch.stack([s0, s1]))
    assert float((lse2 - lse).abs().max()) < 1e-4 and torch.equal(ll2, ll)
    # (5) top-K is sorted, unique, unmasked, and each value is the true score of its id
    crow = torch.arange(0, (M + 1) * 4, 4, device="cuda")
    col = torch.sort(torch.randint(0, N, (M, 4), generator=g, device="cuda"), dim=1).values.reshape(-1)
    vals, ids = ops.topk_eval(U, W, K, crow, col)
    assert torch.all(vals[:, :-1] >= vals[:, 1:])
    true = (U.float().unsqueeze(1) * W[ids.long()].float()).sum(-1)
    assert_rel(vals, true.cpu(), 1e-5, "top-K values")
    assert not torch.any(ids.unsqueeze(-1) == col.view(M, 1, 4))
    # (6) a dense spot check of 8 rows against torch
    S = U[:8].float() @ W.float().T
    S[torch.arange(8, device="cuda").repeat_interleave(4), col[:32]] = -1e23
    rv, ri = torch.sort(S, dim=1, descending=True, stable=True)
    assert torch.equal(ri[:, :K], ids[:8].long()) or float((rv[:, :K] - vals[:8]).abs().max()) < 1e-5


def test_full_size_config4_hstu_retrieval_shard(ops):
    """BASELINE configs[3] per-GPU shard (10M items / 8 = 1.25M, d=256, cosine scores, K=100), B=256 and 4096:
    normalised operands through rb_normalize_rows; sorted, unmasked, true scores; shard split-invariance."""
    g = torch.Generator(device="cuda").manual_seed(9)
    N, d, K = 1_250_000, 256, 100
    W = ops.normalize_rows(torch.randn(N, d, generator=g, device="cuda"), out_dtype=torch.bfloat16)
    for B in (256, 4096):
        U = ops.normalize_rows(torch.randn(B, d, generator=g, device="cuda"), out_dtype=torch.bfloat16)
        crow = torch.arange(0, (B + 1) * 3, 3, device="cuda")
        col = torch.sort(torch.randint(0, N, (B, 3), generator=g, device="cuda"), dim=1).values.reshape(-1)
        vals, ids = ops.topk_eval(U, W, K, crow, col, id_base=0)
        assert torch.all(vals[:, :-1] >= vals[:, 1:]) and torch.all(ids >= 0)
        assert float(vals.max()) <= 1.0 + 1e-2                       # cosine of bf16-rounded unit vectors
        true = (U[:64].float().unsqueeze(1) * W[ids[:64].long()].float()).sum(-1)
        assert_rel(vals[:64], true.cpu(), 1e-5, "top-K values")
        assert not torch.any(ids.unsqueeze(-1) == col.view(B, 1, 3))
        # two half shards + merge == one shard
        h = N // 2
        v0, i0 = ops.topk_eval(U, W[:h], K, crow, col, id_base=0)
        v1, i1 = ops.topk_eval(U, W[h:], K, crow, col, id_base=h)
        mv, mi = ops.topk_merge(torch.stack([v0, v1]), torch.stack([i0, i1]))
        assert torch.equal(mi, ids) and torch.equal(mv, vals)
        # dense spot check of 4 rows
        S = U[:4].float() @ W.float().T
        S[torch.arange(4, device="cuda").repeat_interleave(3), col[:12]] = -1e23
        rv, ri = torch.sort(S, dim=1, descending=True, stable=True)
        assert torch.equal(ri[:, :K], ids[:4].long()) or float((rv[:, :K] - vals[:4]).abs().max()) < 1e-5
        # the CPU oracle in float64 on 24 sampled rows of the FULL-size shard: values, and ids wherever the oracle's
        # neighbouring scores are further apart than the fp32 arithmetic can confuse
        rows = torch.arange(0, B, max(1, B // 24))[:24]
        Sd = orc.score_dense(U[rows].cpu().double(), W.cpu().double())
        sub_crow = torch.arange(0, (len(rows) + 1) * 3, 3)
        sub_col = col.view(B, 3)[rows].reshape(-1).cpu()
        ov, oi = orc.topk_sorted(orc.mask_seen(Sd, sub_crow, sub_col), K + 1)
        assert float((vals[rows].cpu().double() - ov[:, :K]).abs().max()) <= 2e-6
        gap = (ov[:, :-1] - ov[:, 1:]) > 4e-6
        decided = gap[:, :K] & torch.cat([torch.ones_like(gap[:, :1]), gap[:, :K - 1]], 1)
        assert torch.equal(ids[rows].cpu().long()[decided], oi[:, :K][decided])


def test_full_size_config5_bert4rec_shard(ops):
    """BASELINE configs[4] per-GPU shard (50M items / 8 = 6.25M rows + 2 pad columns, d=128, bias head,
    M=4096 masked rows): no logit materialisation (the (M,N) fp32 matrix would be 102 GB), loss / gradient
    properties, label offset of the pad columns."""
    g = torch.Generator(device="cuda").manual_seed(10)
    M, N, d = 4096, 6_250_002, 128
    U = (torch.randn(M, d, generator=g, device="cuda") / d ** 0.25).bfloat16()
    W = (torch.randn(N, d, generator=g, device="cuda") / d ** 0.25).bfloat16()
    bias = torch.randn(N, generator=g, device="cuda") * 0.2
    lab = torch.randint(2, N, (M,), generator=g, device="cuda")      # labels live in [2, N+2) of the N+2 columns
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    m, l, ll, du_un = ops.ce_rowstats(U, W, lab, bias=bias, want_dU=True)
    lse = m + torch.log(l)
    direct = (U.float() * W[lab].float()).sum(-1) + bias[lab]
    assert_rel(ll, direct.cpu(), 1e-5, "label logit")
    assert torch.all(lse >= m) and torch.all(lse <= m + np.log(N) + float(bias.max()) + 1e-3)
    _, dW, db = ops.ce_backward(U, W, lab, lse, 1.0 / M, bias=bias, need_dU=False, need_dW=True, need_dbias=True)
    peak = torch.cuda.max_memory_allocated() - base   # before the checks below allocate anything large
    # rows of (softmax - onehot) sum to zero: sum_j dbias_j == 0 and sum_j dW_j == sum_i (sum_j G_ij) u_i == 0
    assert abs(float(db.double().sum())) <= 1e-3 * float(db.double().abs().sum())
    col_sum = dW.sum(0, dtype=torch.float64)
    assert float(col_sum.abs().max()) <= 5e-3 * float(torch.linalg.vector_norm(dW, ord=1, dim=0, dtype=torch.float64).max())
    # dbias of a label column: softmax mass minus the label count / M
    j = int(lab[0])
    pj = torch.exp((U.float() @ W[j].float()) + bias[j] - lse).sum() / M - float((lab == j).sum()) / M
    assert abs(float(db[j]) - float(pj)) <= 1e-5 + 2e-3 * abs(float(pj))
    # the CPU oracle in float64 at FULL size: lse / loss of 16 sampled query rows (each a 6.25M-term log-sum-exp), and
    # dW / dbias of 8 sampled item rows (each summed over all 4096 query rows against the oracle's global lse)
    rows = torch.arange(0, M, M // 16)[:16]
    Wd, bd = W.cpu().double(), bias.cpu().double()
    Sd = U[rows].cpu().double() @ Wd.T + bd
    lse_ref = torch.logsumexp(Sd, dim=1)
    assert float((lse[rows].cpu().double() - lse_ref).abs().max()) <= 1e-5 * float(lse_ref.abs().max())
    items = torch.cat([lab[:4].cpu(), torch.randint(0, N, (4,), generator=torch.Generator().manual_seed(1))])
    Ud = U.cpu().double()
    P = torch.exp(Ud @ Wd[items].T + bd[items] - lse.cpu().double().unsqueeze(1))           # (M, 8) softmax columns
    onehot = (lab.cpu().unsqueeze(1) == items.unsqueeze(0)).double()
    G = (P - onehot) / M
    assert_grad_bf16(dW[items.to(dW.device)], (G.T @ Ud).float(), "dW rows at full size")
    db_ref = G.sum(0)
    assert float((db[items.to(db.device)].cpu().double() - db_ref).abs().max()) <= 2e-3 * float(db_ref.abs().max()) + 1e-9
    assert peak < 6 * 2 ** 30, f"peak extra memory {peak / 2**30:.1f} GiB: the logits must never be materialised"


@pytest.mark.parametrize("M,N,d", [(513, 4099, 128), (3013, 12101, 64)])
def test_ce_dw_scattered_hot_label(ops, M, N, d):
    """One-hot correction when the rows that share a label are scattered through the batch (every third row,
    as a Zipf-head label is at the bench shape): the owner's warp meets its matches a few per 32-row chunk, so its
    index queue fills across chunks and is flushed with a remainder carried over.  Checked against the oracle
    gradient (whose max-norm these label rows dominate) and, bit for bit, against the fp32 path rounded to bf16."""
    g = torch.Generator().manual_seed(M * 3 + N + d)
    U = bf16_round(torch.randn(M, d, generator=g) * 1.5 / d ** 0.25)
    W = bf16_round(torch.randn(N, d, generator=g) * 1.5 / d ** 0.25)
    lab = torch.randint(0, N, (M,), generator=g)
    lab[1::3] = 11          # scattered hot label whose owner (row 1) is not chunk-aligned
    lab[5::7] = N - 2       # a second one, in the last (partial) item tile
    ref_loss, rdU, rdW, _ = orc.ce_fwd_bwd(U, W, lab)
    Ud, Wd = dev(U).bfloat16().requires_grad_(True), dev(W).bfloat16().requires_grad_(True)
    loss = ops.fused_ce(Ud, Wd, dev(lab))
    loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    assert_grad_bf16(Wd.grad, rdW, "dW")
    Ub, Wb, labd = dev(U).bfloat16(), dev(W).bfloat16(), dev(lab)
    m, l, ll = ops.ce_rowstats(Ub, Wb, labd)
    lse = m + torch.log(l)
    _, dW32, _ = ops.ce_backward(Ub, Wb, labd, lse, 1.0 / M, need_dU=False, need_dW=True)
    _, dWb, _ = ops.ce_backward(Ub, Wb, labd, lse, 1.0 / M, need_dU=False, need_dW=True, dw_dtype=torch.bfloat16)
    assert torch.equal(dWb, dW32.bfloat16())
