"""Multi-GPU parity check (run under torchrun on a GPU box; not collected by pytest): the row-sharded path over
NCCL against the CPU ORACLE in float64 (oracle/reference_path.py on the same bf16-rounded inputs, computed on rank 0
and broadcast) -- loss, the all-reduced dU, every rank's local dW / dbias shard, and the merged masked top-100.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py

Covers the CE train step (loss, dU, local dW shard) at the config-3 shape scaled down, and the
config-4 shape (HSTU-style cosine retrieval, d=256, top-100, normalised operands)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import reference_path as orc  # noqa: E402  (test infrastructure: the checker)
from recboard_b200 import ops, sharded, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True


def check(name, got, ref, tol):
    global ok
    err = float((got.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-30))
    good = err <= tol
    ok &= good
    if rank == 0 or not good:
        print(f"[rank {rank}] {name}: max err / max|ref| = {err:.3e} (tol {tol:.0e}) {'ok' if good else 'FAIL'}", flush=True)


def from_rank0(fn, shapes_dtypes):
    """Run ``fn`` (CPU oracle, float64) on rank 0 only and broadcast its fp32 results."""
    outs = [torch.empty(shape, dtype=dt, device=dev) for shape, dt in shapes_dtypes]
    if rank == 0:
        for o, r in zip(outs, fn()):
            o.copy_(r.to(o.dtype))
    for o in outs:
        dist.broadcast(o, 0)
    return outs


# ---- CE train step, bf16, M=1024, N=200k (+37 to make the shards ragged), d=128
g = torch.Generator(device=dev).manual_seed(11)
M, N, d = 1024, 200_037, 128
U = synth.embeddings(M, d, g, dev, torch.bfloat16, gain=1.5)
W = synth.embeddings(N, d, g, dev, torch.bfloat16, gain=1.5)
labels = synth.zipf_ids(M, N, g, dev)
bias = (torch.randn(N, device=dev, generator=g) * 0.2)
lo, hi = sharded.shard_bounds(N, world, rank)


def oracle_ce(with_bias):
    torch.set_num_threads(os.cpu_count() or 1)
    loss, dU, dW, db = orc.ce_fwd_bwd(U.cpu().double(), W.cpu().double(), labels.cpu(), bias.cpu().double() if with_bias else None)
    return [loss.reshape(1), dU, dW] + ([db] if with_bias else [])


ref_loss, ref_dU, ref_dW = from_rank0(lambda: oracle_ce(False), [((1,), torch.float32), ((M, d), torch.float32), ((N, d), torch.float32)])
Us, Ws = U.clone().requires_grad_(True), W[lo:hi].clone().requires_grad_(True)
loss = sharded.sharded_fused_ce(Us, Ws, labels, lo)
loss.backward()
check("CE loss vs fp64 oracle", loss.detach().reshape(1), ref_loss, 1e-5)
check("CE dU vs fp64 oracle", Us.grad, ref_dU, 2e-3 + 2.0 ** -8)    # north_star bf16 tolerance + the bf16 storage of the gradient
check("CE dW shard vs fp64 oracle", Ws.grad, ref_dW[lo:hi], 2e-3 + 2.0 ** -8)

# the parameter-with-pad-row route: the rank's whole (n_shard + 1, d) parameter, dW added into its existing gradient
Wp = torch.cat([torch.zeros(1, d, dtype=torch.bfloat16, device=dev), W[lo:hi]]).requires_grad_(True)
Wp.grad = torch.zeros_like(Wp)
Us1 = U.clone().requires_grad_(True)
loss1 = sharded.sharded_fused_ce(Us1, Wp, labels, lo, n_skip=1, accumulate=True)
loss1.backward()
check("CE (n_skip, accumulate) loss", loss1.detach().reshape(1), ref_loss, 1e-5)
check("CE (n_skip, accumulate) dW shard", Wp.grad[1:], ref_dW[lo:hi], 2e-3 + 2.0 ** -8)
ok &= bool((Wp.grad[0] == 0).all())

# ---- input-side gather with GLOBAL ids over the sharded table (self.Item.embeddings(seqs), SASRec/main.py:183) and its backward
Bq, Sq = 64, 20
seqs = synth.sequences(Bq, Sq, N, g, dev)                 # 0 = padding, items 1..N
gout = synth.embeddings(Bq * Sq, d, g, dev, torch.bfloat16, gain=0.05).view(Bq, Sq, d)
Wg = W[lo:hi].clone().requires_grad_(True)
emb = sharded.sharded_gather_rows(Wg, seqs - 1, lo, padding_idx=-1)
emb.backward(gout)
ref_emb = orc.gather_rows(torch.cat([torch.zeros(1, d), W.cpu().float()]), seqs.cpu())
ref_gW = orc.scatter_add_rows(gout.cpu().float(), seqs.cpu(), N + 1, padding_idx=0)[1:]
ok &= bool(torch.equal(emb.float().cpu(), ref_emb))
check("sharded gather backward shard", Wg.grad, ref_gW[lo:hi].to(dev), 2.0 ** -8)
if rank == 0:
    print(f"[rank 0] sharded gather forward bit-exact: {bool(torch.equal(emb.float().cpu(), ref_emb))}", flush=True)

# the same through peer memory (CUDA IPC over NVLink): every rank reads the rows where they live, no collective; the
# shard is the rank's whole parameter (one pad row in front), the gradient is added into its existing buffer
Wq = torch.cat([torch.zeros(1, d, dtype=torch.bfloat16, device=dev), W[lo:hi]]).requires_grad_(True)
Wq.grad = torch.zeros_like(Wq)
peers = sharded.PeerTable(Wq, lo, n_skip=1)
emb_p = peers.gather(seqs - 1, padding_idx=-1, accumulate=True)
emb_p.backward(gout)
ok_fwd = bool(torch.equal(emb_p.float().cpu(), ref_emb))
ok &= ok_fwd
check("peer-memory gather backward shard", Wq.grad[1:], ref_gW[lo:hi].to(dev), 2.0 ** -8)
ok &= bool((Wq.grad[0] == 0).all())
with torch.no_grad():      # an owner's update becomes visible to its peers after fence()
    Wq[1:] += 1.0
peers.fence()
emb_p2 = peers.gather(seqs - 1)
ref2 = orc.gather_rows(torch.cat([torch.zeros(1, d), (W.float() + 1.0).bfloat16().float().cpu()]), seqs.cpu())
ok_upd = bool(torch.equal(emb_p2.float().cpu(), ref2))
ok &= ok_upd
peers.close()
if rank == 0:
    print(f"[rank 0] peer-memory gather forward bit-exact: {ok_fwd}; after an owner-side update + fence: {ok_upd}", flush=True)

# ---- BERT4Rec-style bias head (config 5 shape scaled down): bias shard + dbias shard
ref_lb, ref_dUb, ref_dWb, ref_db = from_rank0(lambda: oracle_ce(True), [((1,), torch.float32), ((M, d), torch.float32),
                                                                         ((N, d), torch.float32), ((N,), torch.float32)])
Us2, Ws2, bs2 = U.clone().requires_grad_(True), W[lo:hi].clone().requires_grad_(True), bias[lo:hi].clone().requires_grad_(True)
loss_b = sharded.sharded_fused_ce(Us2, Ws2, labels, lo, bias_shard=bs2)
loss_b.backward()
check("CE+bias loss vs fp64 oracle", loss_b.detach().reshape(1), ref_lb, 1e-5)
check("CE+bias dU vs fp64 oracle", Us2.grad, ref_dUb, 2e-3 + 2.0 ** -8)
check("CE+bias dW shard vs fp64 oracle", Ws2.grad, ref_dWb[lo:hi], 2e-3 + 2.0 ** -8)
check("CE+bias dbias shard vs fp64 oracle", bs2.grad, ref_db[lo:hi], 2e-3)

# ---- HSTU-style retrieval: d=256, K=100, cosine scores, sharded table
B, N2, d2, K = 512, 300_011, 256, 100
Uq = ops.normalize_rows(synth.embeddings(B, d2, g, dev, torch.float32), out_dtype=torch.bfloat16)
W2 = ops.normalize_rows(synth.embeddings(N2, d2, g, dev, torch.float32), out_dtype=torch.bfloat16)
crow, col = synth.seen_csr(B, N2, g, dev)


def oracle_topk():
    S = orc.mask_seen(orc.score_dense(Uq.cpu().double(), W2.cpu().double()), crow.cpu(), col.cpu())
    v, i = orc.topk_sorted(S, K + 1)
    return [v, i]


rv, ri = from_rank0(oracle_topk, [((B, K + 1), torch.float32), ((B, K + 1), torch.int64)])
lo2, hi2 = sharded.shard_bounds(N2, world, rank)
sv, si = sharded.sharded_topk(Uq, W2[lo2:hi2].contiguous(), K, lo2, crow, col)
check("top-100 values vs fp64 oracle", sv, rv[:, :K], 2e-6)
gap_ok = (rv[:, :-1] - rv[:, 1:]) > 4e-6                           # neighbours the fp32 arithmetic can tell apart
agree = (si.long() == ri[:, :K]) | ~gap_ok[:, :K] | ~torch.cat([torch.ones_like(gap_ok[:, :1]), gap_ok[:, :K - 1]], 1)
same = float(agree.float().mean())
if rank == 0:
    print(f"[rank 0] top-100 ids identical wherever the oracle's neighbours are > 4e-6 apart: {same:.6f} "
          f"(exactly identical: {float((si.long() == ri[:, :K]).float().mean()):.6f})", flush=True)
ok &= same == 1.0

flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if float(flag) == 1.0 else "FAIL", f"world={world}", flush=True)
dist.destroy_process_group()
sys.exit(0 if float(flag) == 1.0 else 1)
