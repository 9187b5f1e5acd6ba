"""world_size-2 gloo test (CPU) of the multi-GPU exchange steps: the all-gather of CE row stats and
of per-shard top-K lists, with the oracle standing in for the per-shard kernels."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import reference_path as orc
from recboard_b200 import sharded


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(11)  # same inputs on every rank
    M, N, d, K = 24, 301, 16, 7
    U, W = torch.randn(M, d, generator=g), torch.randn(N, d, generator=g)
    labels = torch.randint(0, N, (M,), generator=g)
    a, b = sharded.shard_bounds(N, world, rank)
    # per-shard partials (what rb_ce_fwd returns on a GPU), here from the oracle
    S = orc.score_dense(U, W[a:b])
    m = S.max(1).values
    l = torch.exp(S - m[:, None]).sum(1)
    inside = (labels >= a) & (labels < b)
    ll = torch.where(inside, S.gather(1, (labels - a).clamp(0, b - a - 1)[:, None]).squeeze(1), torch.zeros(M))
    lse, llg = sharded.merge_rowstats(sharded.allgather_rowstats(m, l, ll))
    loss = (lse - llg).mean()
    # per-shard top-K with global ids, one all-gather, merged by the oracle's merge
    v, i = orc.topk_sorted(S, K)
    av, ai = sharded.allgather_topk(v, (i + a).int())
    mv, mi = orc.merge_topk([(av[r], ai[r]) for r in range(world)], K)
    q.put((rank, float(loss), mv.numpy(), mi.numpy()))   # plain arrays: nothing fd-shared outlives the worker
    dist.destroy_process_group()


def test_two_rank_merges_match_unsharded():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(11)
    M, N, d, K = 24, 301, 16, 7
    U, W = torch.randn(M, d, generator=g), torch.randn(N, d, generator=g)
    labels = torch.randint(0, N, (M,), generator=g)
    ref_loss = orc.ce_loss(U, W, labels)
    rv, ri = orc.topk_sorted(orc.score_dense(U, W), K)
    for rank, loss, mv, mi in outs:
        mv, mi = torch.from_numpy(mv), torch.from_numpy(mi)
        assert abs(loss - float(ref_loss)) < 1e-5
        assert torch.equal(mi, ri) and torch.equal(mv, rv)


# ---- the autograd plumbing of sharded.py (all-gather of stats, all-reduce of dU, local dW / dbias shard,
#      all-gather + merge of top-K) with oracle stand-ins for the per-shard kernels
def _install_oracle_ops():
    from recboard_b200 import ops

    def ce_rowstats(U, W, labels, bias=None, scale=1.0, label_base=0, precision=None, want_dU=False):
        S = orc.score_dense(U, W, bias, scale)
        m = S.max(1).values
        l = torch.exp(S - m[:, None]).sum(1)
        loc = labels - label_base
        inside = (loc >= 0) & (loc < W.shape[0])
        ll = torch.where(inside, S.gather(1, loc.clamp(0, W.shape[0] - 1)[:, None]).squeeze(1), torch.zeros_like(m))
        return m, l, ll

    def ce_backward(U, W, labels, lse, grad_scale, bias=None, scale=1.0, label_base=0, need_dU=True, need_dW=True,
                    need_dbias=False, precision=None, grad_scale_dev=None, dw_dtype=None, dw_out=None, accumulate=False,
                    n_valid=None):
        S = orc.score_dense(U, W, bias, scale)
        G = torch.exp(S - lse[:, None])
        loc = labels - label_base
        inside = (loc >= 0) & (loc < W.shape[0])
        G[torch.arange(len(U))[inside], loc[inside]] -= 1.0
        G = G * grad_scale * (float(grad_scale_dev) if grad_scale_dev is not None else 1.0)
        dW = scale * G.T @ U.float() if need_dW else None
        if dW is not None and dw_out is not None:
            dw_out.copy_(dW)
            dW = dw_out
        return (scale * G @ W.float() if need_dU else None, dW, G.sum(0) if need_dbias else None)

    def topk_eval(U, W, K, seen_crow=None, seen_col=None, bias=None, scale=1.0, id_base=0, precision=None):
        S = orc.score_dense(U, W, bias, scale)
        if seen_crow is not None:
            rows = torch.repeat_interleave(torch.arange(len(U)), seen_crow[1:] - seen_crow[:-1])
            loc = seen_col - id_base
            ok = (loc >= 0) & (loc < W.shape[0])
            S[rows[ok], loc[ok]] = orc.MASK_VALUE
        v, i = orc.topk_sorted(S, K)
        return v, (i + id_base).int()

    ops.fused_du_supported = lambda U, precision, scale: False
    ops.ce_rowstats, ops.ce_backward, ops.topk_eval = ce_rowstats, ce_backward, topk_eval
    ops.topk_merge = lambda av, ai: orc.merge_topk([(av[r], ai[r]) for r in range(av.shape[0])], av.shape[2])


def _inputs():
    g = torch.Generator().manual_seed(12)
    M, N, d, K = 20, 203, 16, 9
    U, W = torch.randn(M, d, generator=g), torch.randn(N, d, generator=g)
    bias = torch.randn(N, generator=g) * 0.3
    labels = torch.randint(0, N, (M,), generator=g)
    seen = [torch.randperm(N, generator=g)[:5].tolist() for _ in range(M)]
    return M, N, d, K, U, W, bias, labels, seen


def _autograd_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _install_oracle_ops()
    M, N, d, K, U, W, bias, labels, seen = _inputs()
    a, b = sharded.shard_bounds(N, world, rank)
    Us, Ws, bs = U.clone().requires_grad_(True), W[a:b].clone().requires_grad_(True), bias[a:b].clone().requires_grad_(True)
    loss = sharded.sharded_fused_ce(Us, Ws, labels, a, bias_shard=bs, scale=0.8)
    (loss * 3.0).backward()
    crow, col = orc.lists_to_csr(seen)
    v, i = sharded.sharded_topk(U, W[a:b], K, a, crow, col, bias_shard=bias[a:b], scale=0.8)
    q.put((rank, a, b, float(loss), *(t.detach().numpy() for t in (Us.grad, Ws.grad, bs.grad, v, i))))
    dist.destroy_process_group()


def test_two_rank_sharded_autograd_and_topk():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_autograd_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    M, N, d, K, U, W, bias, labels, seen = _inputs()
    ref_loss, rdU, rdW, rdb = orc.ce_fwd_bwd(U, W, labels, bias, scale=0.8, grad_out=3.0)
    crow, col = orc.lists_to_csr(seen)
    rv, ri = orc.topk_sorted(orc.mask_seen(orc.score_dense(U, W, bias, 0.8), crow, col), K)
    for rank, a, b, loss, dU, dW, db, v, i in outs:
        dU, dW, db, v, i = (torch.from_numpy(x) for x in (dU, dW, db, v, i))
        assert abs(loss - float(ref_loss)) < 1e-5
        assert torch.allclose(dU, rdU, rtol=1e-4, atol=1e-6)          # the all-reduced full gradient on every rank
        assert torch.allclose(dW, rdW[a:b], rtol=1e-4, atol=1e-6)     # the local shard, never communicated
        assert torch.allclose(db, rdb[a:b], rtol=1e-4, atol=1e-6)
        assert torch.equal(i.long(), ri) and torch.allclose(v, rv)


# ---- input-side gather over a sharded table (global ids, one all-reduce, local backward) and the checkpoint contract
def _gather_ckpt_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from recboard_b200 import ops

    def gather_rows_raw(table, idx):   # what rb_gather_rows does: ids outside [0, rows) give zero rows
        ok = (idx >= 0) & (idx < table.shape[0])
        return torch.where(ok[..., None], table[idx.clamp(0, table.shape[0] - 1)], torch.zeros((), dtype=table.dtype))

    def scatter_add_rows_(grad_table, grad_out, idx, padding_idx=-1):
        ok = (idx >= 0) & (idx < grad_table.shape[0]) & (idx != padding_idx)
        grad_table.index_add_(0, idx[ok], grad_out[ok].to(grad_table.dtype))
        return grad_table

    ops.gather_rows_raw, ops.scatter_add_rows_ = gather_rows_raw, scatter_add_rows_
    g = torch.Generator().manual_seed(31)
    N, P, d, B, S = 101, 1, 8, 5, 7
    full = torch.randn(N, d, generator=g)                       # item rows (global ids 0..N-1 are items, -1 = padding)
    idx = torch.randint(-1, N, (B, S), generator=g)
    gout = torch.randn(B, S, d, generator=g)
    a, b = sharded.shard_bounds(N, world, rank)
    shard = full[a:b].clone().requires_grad_(True)
    out = sharded.sharded_gather_rows(shard, idx, a, padding_idx=-1)
    out.backward(gout)
    # checkpoint round trip: shards -> the reference's single key (with its pad row) -> shards
    sd = sharded.sharded_state_dict({"other": torch.ones(1)}, "Item.embeddings.weight", shard.detach(), N, n_pads=P)
    back = sharded.load_table_shard(sd, "Item.embeddings.weight", world, rank, n_pads=P)
    q.put((rank, a, b, out.detach().numpy(), shard.grad.numpy(), sd["Item.embeddings.weight"].numpy(), back.numpy()))
    dist.destroy_process_group()


def test_two_rank_sharded_gather_and_checkpoint():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_ckpt_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(31)
    N, P, d, B, S = 101, 1, 8, 5, 7
    full = torch.randn(N, d, generator=g).requires_grad_(True)
    idx = torch.randint(-1, N, (B, S), generator=g)
    gout = torch.randn(B, S, d, generator=g)
    table = torch.cat([torch.zeros(P, d), full])                 # the reference's layout: pad row first
    ref = torch.nn.functional.embedding(idx + P, table, padding_idx=0)
    ref.backward(gout)
    for rank, a, b, out, grad, ckpt, back in outs:
        assert torch.allclose(torch.from_numpy(out), ref.detach())
        assert torch.allclose(torch.from_numpy(grad), full.grad[a:b], atol=1e-6)
        assert torch.equal(torch.from_numpy(ckpt)[P:], full.detach()) and bool((torch.from_numpy(ckpt)[:P] == 0).all())
        assert torch.equal(torch.from_numpy(back), full.detach()[a:b])
