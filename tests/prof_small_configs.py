"""Timings of BASELINE configs[0] and configs[1] in fp32-parity mode (CUDA events, one GPU):
  config 1: SASRec full-catalog CE train step at the Beauty shape (M=3013 query rows, N=12101 items, d=64)
            and its full-ranking evaluation batch (B=512, top-50);
  config 2: MF-BPR / LightGCN full-ranking top-20 with seen masking at the Yelp2018 shape (31668 users,
            38048 items, d=64), all users in one call."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from recboard_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(2027)
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


M, N, d = 3013, 12101, 64
U = synth.embeddings(M, d, g, dev).requires_grad_(True)
W = synth.embeddings(N, d, g, dev).requires_grad_(True)
lab = synth.zipf_ids(M, N, g, dev)


def train_step():
    U.grad = W.grad = None
    ops.fused_ce(U, W, lab, precision="fp32").backward()


ms = t(train_step)
out["config1 CE train step fp32 (3013 x 12101, d=64)"] = {"ms": round(ms, 4), "pairs_per_s": M * N / ms * 1e3}
Ue = synth.embeddings(512, d, g, dev)
crow, col = synth.seen_csr(512, N, g, dev)
ms = t(lambda: ops.topk_eval(Ue, W.detach(), 50, crow, col, precision="fp32"))
out["config1 eval batch fp32 (512 x 12101, top-50, seen mask)"] = {"ms": round(ms, 4), "pairs_per_s": 512 * N / ms * 1e3}

B2, N2 = 31668, 38048
U2 = synth.embeddings(B2, d, g, dev); W2 = synth.embeddings(N2, d, g, dev)
crow2, col2 = synth.seen_csr(B2, N2, g, dev, mean_len=40.0)
ms = t(lambda: ops.topk_eval(U2, W2, 20, crow2, col2, precision="fp32"), 5)
out["config2 top-20 fp32, all 31668 users x 38048 items in one call"] = {
    "ms": round(ms, 4), "pairs_per_s": B2 * N2 / ms * 1e3, "algorithmic_tflops_fp32": round(2.0 * B2 * N2 * d / ms / 1e9, 1)}
ms = t(lambda: ops.topk_eval(U2.bfloat16(), W2.bfloat16(), 20, crow2, col2), 5)
out["config2 top-20 bf16 (same shapes)"] = {"ms": round(ms, 4), "pairs_per_s": B2 * N2 / ms * 1e3}
print(json.dumps(out, indent=1))
