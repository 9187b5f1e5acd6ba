import sys, json
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from recboard_b200 import ops, synth  # noqa: E402
dev = torch.device("cuda", 0)
def t(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
out = {}
for d in (32, 64, 128):
    g = torch.Generator(device=dev).manual_seed(d)
    M, N = 4096, 1_000_000
    W = synth.embeddings(N, d, g, dev, torch.bfloat16, gain=1.5)
    U = synth.embeddings(M, d, g, dev, torch.bfloat16, gain=1.5)
    lab = synth.zipf_ids(M, N, g, dev)
    m, l, ll = ops.ce_rowstats(U, W, lab)
    lse = m + torch.log(l)
    out[f"d{d}"] = {"fwd_dU_ms": round(t(lambda: ops.ce_rowstats(U, W, lab, want_dU=True)), 4),
                    "dW_ms": round(t(lambda: ops.ce_backward(U, W, lab, lse, 1.0 / M, need_dU=False, need_dW=True, dw_dtype=torch.bfloat16)), 4)}
print(json.dumps(out))
