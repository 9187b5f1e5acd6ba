"""Shape survey: fused CE train step (fwd + dU + dW through autograd) and masked top-K over a grid of shapes, to spot
shapes that fall off the fast paths.  One line per shape: ms and algorithmic TFLOP/s (6 M N d / 2 M N d)."""
import sys, json
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from recboard_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda", 0)

def t(fn, reps=8):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

rows = []
for M, N, d in [(128, 1_000_000, 128), (512, 1_000_000, 128), (2048, 1_000_000, 128), (4096, 1_000_000, 64), (4096, 1_000_000, 32),
                (4096, 200_000, 128), (16384, 1_000_000, 128), (40960, 1_000_000, 128), (4096, 1_000_003, 72)]:
    g = torch.Generator(device=dev).manual_seed(M + d)
    W = synth.embeddings(N, d, g, dev, torch.bfloat16, gain=1.5).requires_grad_(True)
    U = synth.embeddings(M, d, g, dev, torch.bfloat16, gain=1.5).requires_grad_(True)
    lab = synth.zipf_ids(M, N, g, dev)
    crow, col = synth.seen_csr(M, N, g, dev)

    def train():
        U.grad = None; W.grad = None
        ops.fused_ce(U, W, lab).backward()

    tt = t(train)
    te = t(lambda: ops.topk_eval(U.detach(), W.detach(), 50, crow, col))
    rows.append({"M": M, "N": N, "d": d, "ce_train_ms": round(tt, 4), "ce_tflops": round(6.0 * M * N * d / tt / 1e9, 1),
                 "top50_ms": round(te, 4), "top50_tflops": round(2.0 * M * N * d / te / 1e9, 1)})
    print(json.dumps(rows[-1]), flush=True)
    del W, U
    torch.cuda.empty_cache()
