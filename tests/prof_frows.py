"""Micro-benchmarks of the HBM-bound rows (CUDA events; algorithmic bytes per DESIGN.md 3.3):
gather-dot forward/backward (f-1), CSR SpMM (f-3), row normalisation (a11), gather / scatter-add (a2, a3)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from recboard_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
out = {}


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def record(name, ms, nbytes, **kw):
    out[name] = {"ms": round(ms, 4), "algorithmic_GB": round(nbytes / 1e9, 4), "GBps": round(nbytes / ms / 1e6, 1), **kw}


# ---- gather-dot: sampled softmax (HSTU shape: 1 + 512 candidates) on a 1M x 128 bf16 table
M, K, N, d = 4096, 513, 1_000_000, 128
U = synth.embeddings(M, d, g, dev, torch.bfloat16).requires_grad_(True)
W = synth.embeddings(N, d, g, dev, torch.bfloat16).requires_grad_(True)
idx = synth.zipf_ids(M * K, N, g, dev).view(M, K)
G = torch.randn(M, K, device=dev, generator=g)
record("gather_dot_fwd bf16 4096x513 of 1M x128", t(lambda: ops.gather_dot(U.detach(), W.detach(), idx)),
       M * K * (8 + d * 2 + 4) + M * d * 2)
S = ops.gather_dot(U, W, idx)
record("gather_dot_bwd (dU + dense dTable)", t(lambda: torch.autograd.grad(S, (U, W), G, retain_graph=True)),
       M * K * (8 + 4 + d * 2) * 2 + M * d * 4 + 2 * N * d * 4, note="includes zero-fill + cast of the dense (N,d) table gradient")
ref_ms = t(lambda: torch.einsum("md,mkd->mk", U.detach(), W.detach()[idx]))
out["gather_dot_fwd torch eager (gather + einsum)"] = {"ms": round(ref_ms, 4)}

# ---- gather-dot at the Beauty shape, fp32 (config 1): 3013 rows x 513 candidates of 12101 x 64
M2, N2, d2 = 3013, 12101, 64
U2 = synth.embeddings(M2, d2, g, dev); W2 = synth.embeddings(N2, d2, g, dev)
idx2 = synth.zipf_ids(M2 * K, N2, g, dev).view(M2, K)
record("gather_dot_fwd fp32 3013x513 of 12101 x64", t(lambda: ops.gather_dot(U2, W2, idx2)), M2 * K * (8 + d2 * 4 + 4) + M2 * d2 * 4)

# ---- SpMM: LightGCN propagation at the Yelp2018 shape (69 716 nodes, 3.1M non-zeros, d = 64)
R, nnz_half, ds = 31_668 + 38_048, 1_561_406, 64
r = torch.randint(0, 31_668, (nnz_half,), device=dev, generator=g)
c = torch.randint(31_668, R, (nnz_half,), device=dev, generator=g)
v = torch.rand(nnz_half, device=dev, generator=g)
A = torch.sparse_coo_tensor(torch.stack([torch.cat([r, c]), torch.cat([c, r])]), torch.cat([v, v]), (R, R)).coalesce().to_sparse_csr()
X = torch.randn(R, ds, device=dev, generator=g)
nnz = A.values().numel()
record("spmm_csr 69716 nodes, %d nnz, d=64" % nnz, t(lambda: ops.spmm_raw(A, X)), nnz * (12 + ds * 4) + 2 * R * ds * 4)
out["spmm torch (A @ X)"] = {"ms": round(t(lambda: A @ X), 4)}

# ---- normalisation of a 2M x 256 fp32 table to bf16 (config-4 shard + a bit)
Wn = torch.randn(2_000_000, 256, device=dev, generator=g)
record("normalize_rows 2M x 256 fp32 -> bf16", t(lambda: ops.normalize_rows(Wn, out_dtype=torch.bfloat16)), Wn.numel() * 6)
out["normalize torch (F.normalize + cast)"] = {"ms": round(t(lambda: torch.nn.functional.normalize(Wn, dim=-1).bfloat16()), 4)}

# ---- gather / scatter-add at the bench shape
table = synth.embeddings(1_000_001, 128, g, dev, torch.bfloat16)
seqs = synth.sequences(4096, 50, 1_000_000, g, dev)
go = synth.embeddings(4096 * 50, 128, g, dev, torch.bfloat16, gain=0.01)
tg = torch.zeros(1_000_001, 128, device=dev)
record("gather_rows 4096x50 of 1M x128 bf16", t(lambda: ops.gather_rows_raw(table, seqs)), 4096 * 50 * (8 + 2 * 128 * 2))
touched = int(torch.unique(seqs).numel())
record("scatter_add_rows 204800 ids (%d distinct rows)" % touched, t(lambda: ops.scatter_add_rows_(tg, go, seqs.view(-1), padding_idx=0)),
       4096 * 50 * 8 + int((seqs != 0).sum()) * 128 * 2 + 2 * touched * 128 * 4)
print(json.dumps(out, indent=1))
