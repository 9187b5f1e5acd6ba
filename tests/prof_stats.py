"""Timing probe: statistics-only CE forward sweep (EPI_LSE), bf16 4096 x 1M x 128 and fp32-parity config-2-like shapes."""
import sys, json, torch
sys.path.insert(0, ".")
from recboard_b200 import ops

def t(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

out = {}
for name, M, N, d, dt in (("bf16_4096x1M_d128", 4096, 1_000_000, 128, torch.bfloat16), ("bf16_4096x1M_d64", 4096, 1_000_000, 64, torch.bfloat16),
                          ("fp32_31668x38048_d64", 31668, 38048, 64, torch.float32), ("fp32_3013x12101_d64", 3013, 12101, 64, torch.float32),
                          ("bf16_bias_4096x1M_d128", 4096, 1_000_000, 128, torch.bfloat16)):
    g = torch.Generator(device="cuda").manual_seed(d)
    U = (torch.randn(M, d, device="cuda", generator=g) / d ** 0.25).to(dt)
    W = (torch.randn(N, d, device="cuda", generator=g) / d ** 0.25).to(dt)
    lab = torch.randint(0, N, (M,), device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g) * 0.1 if "bias" in name else None
    out[name] = t(lambda: ops.ce_rowstats(U, W, lab, bias=bias))
g = torch.Generator(device="cuda").manual_seed(5)
M, N, d = 4096, 1_000_000, 128
U = (torch.randn(M, d, device="cuda", generator=g) / d ** 0.25).bfloat16()
W = (torch.randn(N, d, device="cuda", generator=g) / d ** 0.25).bfloat16()
bias = torch.randn(N, device="cuda", generator=g) * 0.1
out["topk50_bf16_4096x1M_d128"] = t(lambda: ops.topk_eval(U, W, 50))
out["topk50_bias_bf16_4096x1M_d128"] = t(lambda: ops.topk_eval(U, W, 50, bias=bias))
print(json.dumps(out))
