"""Measurement helpers of bench.py that are not the headline step (bench-side code, not part of the product package):

* ``eager_baseline``  -- the reference's own hot-path LINES in PyTorch eager on the same B200 (the reference has no
  kernels of its own: cuBLAS einsum + ``F.cross_entropy`` + autograd, SASRec/main.py:217-219,249; dense seen mask +
  ``scores[seen] = -1e23`` + one ``torch.topk`` per metric@k, UniSRec/main.py:408-435), timed beside the fused path on
  BASELINE.json configs 1, 2 and 3.  This is the bar on the box.
* ``strong_scaling``  -- a FIXED-size 10M-item catalog row-sharded over the ranks (north_star: ">= 6.5x at 8 GPUs on a
  10M-item sharded table"): CE train at d = 128 and the config-4 retrieval (d = 256, L2-normalised, top-100).
* ``config5``         -- BERT4Rec masked-item CE over a 50M-item catalog with bias, row-sharded over the ranks.

Everything is device-timed with CUDA events (max over ranks for the sharded blocks).  PyTorch eager here is a
BASELINE being measured, never a path the product takes.
"""
from __future__ import annotations

from typing import Callable, Dict

import torch
import torch.distributed as dist
import torch.nn.functional as F

from recboard_b200 import ops, sharded, synth

MONITORS = ["HITRATE@1", "HITRATE@5", "HITRATE@10", "HITRATE@20", "HITRATE@50", "NDCG@5", "NDCG@10", "NDCG@20", "NDCG@50"]


def time_ms(fn: Callable, reps: int = 3, warmup: int = 1, world: int = 1) -> float:
    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


# ------------------------------------------------------------------------------------------------ eager reference lines
def _eager_train(U, W, labels, autocast: bool):
    Uq, Wq = U.detach().requires_grad_(True), W.detach().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        logits = torch.einsum("MD,ND->MN", Uq, Wq)                  # SASRec/main.py:217
        loss = F.cross_entropy(logits, labels)                      # :219 (CrossEntropy4Logits)
    loss.backward()                                                 # :249
    return loss


def _eager_eval(U, W, seen_crow, seen_col, tgt_crow, tgt_col, ks, autocast: bool, batch: int):
    """UniSRec/main.py:400-435 per batch of ``batch`` rows: dense scores, dense seen mask, ``scores[seen] = -1e23``,
    dense targets, one torch.topk per metric@k, ``.item()`` per metric."""
    B, N = U.shape[0], W.shape[0]
    out = 0.0
    for lo in range(0, B, batch):
        hi = min(lo + batch, B)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            scores = torch.einsum("BD,ND->BN", U[lo:hi], W).float()  # :408 (model(data, ranking="full"))
        a, b = int(seen_crow[lo]), int(seen_crow[hi])
        seen = torch.sparse_csr_tensor(seen_crow[lo:hi + 1] - a, seen_col[a:b], torch.ones(b - a, device=U.device),
                                       (hi - lo, N)).to_dense().bool()                                       # :410-412
        scores[seen] = -1e23                                                                                  # :413
        a, b = int(tgt_crow[lo]), int(tgt_crow[hi])
        targets = torch.sparse_csr_tensor(tgt_crow[lo:hi + 1] - a, tgt_col[a:b], torch.ones(b - a, device=U.device),
                                          (hi - lo, N)).to_dense()                                           # :414
        for name, k in ks:                                                                                   # :428-435
            idx = torch.topk(scores, k, dim=1).indices
            h = targets.gather(1, idx)
            if name == "HITRATE":
                out += (h.sum(-1) > 0).float().mean().item()
            else:
                w = 1.0 / torch.log2(torch.arange(k, device=U.device, dtype=torch.float32) + 2.0)
                out += ((h * w).sum(-1)).mean().item()
    return out


def _fused_eval(U, W, K, seen_crow, seen_col, tgt_crow, tgt_col, precision, monitors):
    from recboard_b200 import metrics as MX
    _, ids = ops.topk_eval(U, W, K, seen_crow, seen_col, precision=precision)
    return MX.batch_metrics(ids, tgt_crow, tgt_col, W.shape[0], monitors, exact=False)   # one pass, one small read


def eager_baseline(dev: torch.device, reps: int = 3) -> Dict[str, Dict]:
    """PyTorch eager (the reference lines) vs the fused path on configs 1-3, one GPU.  Times in ms, pairs/s from
    rows x items / time; ``speedup_<mode>`` = eager in that precision / fused in that mode."""
    out = {}
    ks9 = [(m.split("@")[0], int(m.split("@")[1])) for m in MONITORS]
    cfgs = {
        # SASRec, Beauty shape: M ~ 3 013 non-pad positions of 512 x 50, N = 12 101, d = 64; eval batch 512, monitors @<=50
        "config1": dict(M=3013, B=512, N=12101, d=64, K=50, ks=ks9, eval_batch=512, fused="fp32"),
        # MF-BPR / LightGCN, Yelp2018 shape: all 31 668 users x 38 048 items, d = 64, top-20 (the reference: batches of 512)
        "config2": dict(M=0, B=31668, N=38048, d=64, K=20, ks=[("HITRATE", 20), ("NDCG", 20)], eval_batch=512, fused="fp32"),
        # SASRec bf16 full softmax: 4096 x 1M x 128
        "config3": dict(M=4096, B=4096, N=1_000_000, d=128, K=50, ks=ks9, eval_batch=4096, fused="bf16"),
    }
    for name, c in cfgs.items():
        g = torch.Generator(device=dev).manual_seed(2026 + int(name[-1]))
        N, d, B, M = c["N"], c["d"], c["B"], c["M"]
        W = synth.embeddings(N, d, g, dev, torch.float32, gain=1.5)
        Ue = synth.embeddings(B, d, g, dev, torch.float32, gain=1.5)
        seen_crow, seen_col = synth.seen_csr(B, N, g, dev)
        tgt = synth.targets(B, N, g, dev, (seen_crow, seen_col))
        tgt_crow = torch.arange(B + 1, device=dev, dtype=torch.int64)
        res = {"shape": {k: c[k] for k in ("M", "B", "N", "d", "K")}}
        modes = ("fp32", "bf16") if c["fused"] == "fp32" else ("bf16",)   # like against like: fp32 parity mode vs eager fp32
        if M:                                                               # (3xTF32 scores, exact fp32 gradients), bf16 mode vs
            Ut = synth.embeddings(M, d, g, dev, torch.float32, gain=1.5)    # eager under bf16 autocast
            labels = synth.zipf_ids(M, N, g, dev)
            eager = {"fp32": time_ms(lambda: _eager_train(Ut, W, labels, False), reps),
                     "bf16": time_ms(lambda: _eager_train(Ut, W, labels, True), reps)}
            tr = {"eager_fp32_ms": eager["fp32"], "eager_bf16_autocast_ms": eager["bf16"]}
            for fp in modes:
                cast = (lambda x: x.bfloat16()) if fp == "bf16" else (lambda x: x)
                Uf, Wf = cast(Ut).requires_grad_(True), cast(W).requires_grad_(True)

                def fused_train():
                    Uf.grad = None; Wf.grad = None
                    ops.fused_ce(Uf, Wf, labels, precision=fp).backward()

                tf = time_ms(fused_train, max(reps, 5), 2)
                tr[f"fused_{fp}_ms"] = tf
                tr[f"speedup_{fp}"] = eager[fp] / tf
                tr[f"fused_{fp}_pairs_per_s"] = M * N / (tf * 1e-3)
            tr["eager_pairs_per_s"] = M * N / (min(eager.values()) * 1e-3)
            res["train"] = tr
        eager = {"fp32": time_ms(lambda: _eager_eval(Ue, W, seen_crow, seen_col, tgt_crow, tgt, c["ks"], False, c["eval_batch"]), 2),
                 "bf16": time_ms(lambda: _eager_eval(Ue, W, seen_crow, seen_col, tgt_crow, tgt, c["ks"], True, c["eval_batch"]), 2)}
        ev = {"eager_fp32_ms": eager["fp32"], "eager_bf16_autocast_ms": eager["bf16"],
              "note": "both sides evaluate the same metric@k list and include its device->host reads"}
        for fp in modes:
            cast = (lambda x: x.bfloat16()) if fp == "bf16" else (lambda x: x)
            Uef, Wef = cast(Ue), cast(W)
            mons = [f"{n}@{k}" for n, k in c["ks"]]   # the same metric@k list the eager side evaluates
            ef = time_ms(lambda: _fused_eval(Uef, Wef, c["K"], seen_crow, seen_col, tgt_crow, tgt, fp, mons), max(reps, 5), 2)
            ev[f"fused_{fp}_ms"] = ef
            ev[f"speedup_{fp}"] = eager[fp] / ef
            ev[f"fused_{fp}_pairs_per_s"] = B * N / (ef * 1e-3)
        ev["eager_pairs_per_s"] = B * N / (min(eager.values()) * 1e-3)
        res["eval"] = ev
        out[name] = res
        del W, Ue
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ fixed-size sharded blocks
def strong_scaling(dev: torch.device, world: int, rank: int, n_total: int = 10_000_000, rows: int = 4096, reps: int = 3) -> Dict:
    """A 10M-item catalog, FIXED size, row-sharded over ``world`` ranks: (i) bf16 CE train step at d = 128
    (fused forward + dU, dW into the parameter's gradient; one all-gather of row statistics, one all-reduce of dU);
    (ii) config 4: cosine retrieval at d = 256, both sides L2-normalised, top-100, one all-gather + merge."""
    a, b = sharded.shard_bounds(n_total, world, rank)
    n_shard = b - a
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    gq = torch.Generator(device=dev).manual_seed(78)      # queries: identical on every rank
    out = {"n_items_total": n_total, "rows": rows, "n_gpus": world}
    # (i) CE train
    table = synth.embeddings(n_shard, 128, g, dev, torch.bfloat16, gain=1.5).requires_grad_(True)
    U = synth.embeddings(rows, 128, gq, dev, torch.bfloat16, gain=1.5)
    labels = synth.zipf_ids(rows, n_total, gq, dev)

    def train():
        table.grad = None                      # optimizer.zero_grad(): the dW pass writes the new gradient buffer
        Uq = U.detach().requires_grad_(True)
        if world > 1:
            loss = sharded.sharded_fused_ce(Uq, table, labels, a, accumulate=True)
        else:
            loss = ops.fused_ce(Uq, table, labels, accumulate=True)
        loss.backward()
        return loss

    t_train = time_ms(train, reps, 2, world)
    out["ce_train"] = {"ms": t_train, "pairs_per_s": rows * n_total / (t_train * 1e-3), "d": 128,
                       "algorithmic_tflops": 6.0 * rows * n_total * 128 / (t_train * 1e-3) / 1e12}
    del table, U
    torch.cuda.empty_cache()
    # (ii) config 4 retrieval
    raw = synth.embeddings(n_shard, 256, g, dev, torch.bfloat16, gain=1.0)
    Wn = ops.normalize_rows(raw, out_dtype=torch.bfloat16)                 # HSTU/main.py:182-184, once per sweep
    del raw
    Un = ops.normalize_rows(synth.embeddings(rows, 256, gq, dev, torch.float32, gain=1.0), out_dtype=torch.bfloat16)
    seen_crow, seen_col = synth.seen_csr(rows, n_total, gq, dev)
    for B in (rows, 256):
        crow = seen_crow[:B + 1].contiguous()
        col = seen_col[:int(crow[-1])].contiguous()
        Ub = Un[:B].contiguous()

        def ev():
            if world > 1:
                return sharded.sharded_topk(Ub, Wn, 100, a, crow, col)
            return ops.topk_eval(Ub, Wn, 100, crow, col)

        t = time_ms(ev, reps, 2, world)
        out[f"eval_top100_B{B}"] = {"ms": t, "pairs_per_s": B * n_total / (t * 1e-3), "d": 256,
                                    "algorithmic_tflops": 2.0 * B * n_total * 256 / (t * 1e-3) / 1e12,
                                    "table_gb_per_s": n_total * 256 * 2 / (t * 1e-3) / 1e9}
    return out


def config5(dev: torch.device, world: int, rank: int, n_total: int = 50_000_000, rows: int = 4096, reps: int = 2) -> Dict:
    """BASELINE configs[4]: BERT4Rec masked-item CE over a 50M-item catalog (+ bias), d = 128, row-sharded; the logits
    (rows x 50M fp32 = 819 GB at 4096 rows) are never formed.  Peak memory is reported per rank."""
    a, b = sharded.shard_bounds(n_total, world, rank)
    n_shard = b - a
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    gq = torch.Generator(device=dev).manual_seed(98)
    torch.cuda.reset_peak_memory_stats(dev)
    W = synth.embeddings(n_shard, 128, g, dev, torch.bfloat16, gain=1.5).requires_grad_(True)
    bias = (torch.randn(n_shard, device=dev, generator=g) * 0.1).requires_grad_(True)
    U = synth.embeddings(rows, 128, gq, dev, torch.bfloat16, gain=1.5)
    labels = synth.zipf_ids(rows, n_total, gq, dev)

    def step():
        W.grad = None; bias.grad = None
        Uq = U.detach().requires_grad_(True)
        if world > 1:
            loss = sharded.sharded_fused_ce(Uq, W, labels, a, bias_shard=bias)
        else:
            loss = ops.fused_ce(Uq, W, labels, bias=bias)
        loss.backward()
        return loss

    t = time_ms(step, reps, 1, world)
    loss = float(step().detach())
    return {"n_items_total": n_total, "rows": rows, "d": 128, "n_gpus": world, "ms": t, "loss": loss,
            "pairs_per_s": rows * n_total / (t * 1e-3), "algorithmic_tflops": 6.0 * rows * n_total * 128 / (t * 1e-3) / 1e12,
            "peak_mem_gib_per_rank": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
