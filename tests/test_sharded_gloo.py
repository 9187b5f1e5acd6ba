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
    q.put((rank, float(loss), mv, mi))
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
        assert abs(loss - float(ref_loss)) < 1e-5
        assert torch.equal(mi, ri) and torch.equal(mv, rv)
