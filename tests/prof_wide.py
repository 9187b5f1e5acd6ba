"""Timing probe: fused CE passes at d = 256 (d-split variant) next to d = 128, 4096 rows x 1M items, bf16."""
import sys, json, torch
sys.path.insert(0, ".")
from recboard_b200 import ops

def t(fn, reps=5):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

out = {}
M, N = 4096, 1_000_000
for d in (128, 192, 256):
    g = torch.Generator(device="cuda").manual_seed(d)
    U = (torch.randn(M, d, device="cuda", generator=g) / d ** 0.25).bfloat16()
    W = (torch.randn(N, d, device="cuda", generator=g) / d ** 0.25).bfloat16()
    lab = torch.randint(0, N, (M,), device="cuda", generator=g)
    m, l, ll = ops.ce_rowstats(U, W, lab)
    lse = m + torch.log(l)
    tf = t(lambda: ops.ce_rowstats(U, W, lab, want_dU=True))
    tw = t(lambda: ops.ce_backward(U, W, lab, lse, 1.0 / M, need_dU=False, need_dW=True, dw_dtype=torch.bfloat16))
    flop = 2.0 * M * N * d
    out[f"d{d}"] = {"fwd_dU_ms": tf, "dW_ms": tw, "algorithmic_tflops": 3 * flop / ((tf + tw) * 1e-3) / 1e12}
print(json.dumps(out))
