"""Probe: time the two fused CE passes (CUDA events) at bench sizes.  RB_SO_SUFFIX selects an experiment build."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from recboard_b200 import ops, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
M, D = 4096, 128
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
U = synth.embeddings(M, D, g, dev, torch.bfloat16, gain=1.5)
W = synth.embeddings(N, D, g, dev, torch.bfloat16, gain=1.5)
labels = synth.zipf_ids(M, N, g, dev)
m, l, ll = ops.ce_rowstats(U, W, labels)
lse = m + torch.log(l)


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


print("fwd+dU ms", t(lambda: ops.ce_rowstats(U, W, labels, want_dU=True)))
print("dW ms", t(lambda: ops.ce_backward(U, W, labels, lse, 1.0 / M, need_dU=False, need_dW=True)))
