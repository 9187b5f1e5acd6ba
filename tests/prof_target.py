"""Profiling target for ncu: one launch of each sweep of the hot path at bench.py's sizes
(M=4096, N=1M, d=128, bf16).  Usage: ncu ... python tests/prof_target.py [N] [which]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from recboard_b200 import ops, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
which = sys.argv[2] if len(sys.argv) > 2 else "fwd,dU,dW,topk"
M, D, K = 4096, 128, 50
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
U = synth.embeddings(M, D, g, dev, torch.bfloat16, gain=1.5)
W = synth.embeddings(N, D, g, dev, torch.bfloat16, gain=1.5)
labels = synth.zipf_ids(M, N, g, dev)
crow, col = synth.seen_csr(M, N, g, dev)
torch.cuda.synchronize()
m, l, ll = ops.ce_rowstats(U, W, labels)
lse = m + torch.log(l)
if "fwd" in which:
    ops.ce_rowstats(U, W, labels)
if "dU" in which:
    ops.ce_backward(U, W, labels, lse, 1.0 / M, need_dU=True, need_dW=False)
if "dW" in which:
    ops.ce_backward(U, W, labels, lse, 1.0 / M, need_dU=False, need_dW=True, dw_dtype=torch.bfloat16)   # as in the bench step
if "topk" in which:
    ops.topk_eval(U, W, K, crow, col)
torch.cuda.synchronize()
print("done")
