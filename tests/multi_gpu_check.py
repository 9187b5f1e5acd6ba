"""Multi-GPU parity check (run under torchrun on a GPU box; not collected by pytest):
the row-sharded path over NCCL must reproduce the single-GPU full-table results.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py

Covers the CE train step (loss, dU, local dW shard) at the config-3 shape scaled down, and the
config-4 shape (HSTU-style cosine retrieval, d=256, top-100, normalised operands)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
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


# ---- CE train step, bf16, M=1024, N=200k (+37 to make the shards ragged), d=128
g = torch.Generator(device=dev).manual_seed(11)
M, N, d = 1024, 200_037, 128
U = synth.embeddings(M, d, g, dev, torch.bfloat16, gain=1.5)
W = synth.embeddings(N, d, g, dev, torch.bfloat16, gain=1.5)
labels = synth.zipf_ids(M, N, g, dev)
Uf, Wf = U.clone().requires_grad_(True), W.clone().requires_grad_(True)
ref_loss = ops.fused_ce(Uf, Wf, labels)
ref_loss.backward()
lo, hi = sharded.shard_bounds(N, world, rank)
Us, Ws = U.clone().requires_grad_(True), W[lo:hi].clone().requires_grad_(True)
loss = sharded.sharded_fused_ce(Us, Ws, labels, lo)
loss.backward()
check("CE loss", loss.detach().reshape(1), ref_loss.detach().reshape(1), 1e-5)
check("CE dU", Us.grad, Uf.grad, 2.0 ** -7)   # both sides are bf16 tensors: one ulp of the largest element is 2^-8
check("CE dW shard", Ws.grad, Wf.grad[lo:hi], 2.0 ** -7)

# ---- BERT4Rec-style bias head (config 5 shape scaled down): bias shard + dbias shard
bias = (torch.randn(N, device=dev, generator=g) * 0.2)
Ub, Wb, bb = U.clone().requires_grad_(True), W.clone().requires_grad_(True), bias.clone().requires_grad_(True)
ref_b = ops.fused_ce(Ub, Wb, labels, bias=bb)
ref_b.backward()
Us2, Ws2, bs2 = U.clone().requires_grad_(True), W[lo:hi].clone().requires_grad_(True), bias[lo:hi].clone().requires_grad_(True)
loss_b = sharded.sharded_fused_ce(Us2, Ws2, labels, lo, bias_shard=bs2)
loss_b.backward()
check("CE+bias loss", loss_b.detach().reshape(1), ref_b.detach().reshape(1), 1e-5)
check("CE+bias dU", Us2.grad, Ub.grad, 2.0 ** -7)
check("CE+bias dW shard", Ws2.grad, Wb.grad[lo:hi], 2.0 ** -7)
check("CE+bias dbias shard", bs2.grad, bb.grad[lo:hi], 1e-4)

# ---- HSTU-style retrieval: d=256, K=100, cosine scores, sharded table
B, N2, d2, K = 512, 300_011, 256, 100
Uq = ops.normalize_rows(synth.embeddings(B, d2, g, dev, torch.float32), out_dtype=torch.bfloat16)
W2 = ops.normalize_rows(synth.embeddings(N2, d2, g, dev, torch.float32), out_dtype=torch.bfloat16)
crow, col = synth.seen_csr(B, N2, g, dev)
rv, ri = ops.topk_eval(Uq, W2, K, crow, col)
lo2, hi2 = sharded.shard_bounds(N2, world, rank)
sv, si = sharded.sharded_topk(Uq, W2[lo2:hi2].contiguous(), K, lo2, crow, col)
check("top-100 values", sv, rv, 1e-6)
same = float((si == ri).float().mean())
if rank == 0:
    print(f"[rank 0] top-100 ids identical: {same:.6f}", flush=True)
ok &= same == 1.0

flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if float(flag) == 1.0 else "FAIL", f"world={world}", flush=True)
dist.destroy_process_group()
sys.exit(0 if float(flag) == 1.0 else 1)
