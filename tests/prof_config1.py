"""Profiling target: one fused CE train step at the config-1 shape (fp32 parity mode), for an ncu launch list."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from recboard_b200 import ops, synth  # noqa: E402

M, N, d = 3013, 12101, 64
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
U = synth.embeddings(M, d, g, dev, torch.float32, gain=1.5).requires_grad_(True)
W = synth.embeddings(N, d, g, dev, torch.float32, gain=1.5).requires_grad_(True)
labels = synth.zipf_ids(M, N, g, dev)
for _ in range(3):
    U.grad = None; W.grad = None
    ops.fused_ce(U, W, labels, precision="fp32").backward()
torch.cuda.synchronize()
if "--time" in sys.argv:
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        U.grad = None; W.grad = None
        ops.fused_ce(U, W, labels, precision="fp32").backward()
    b.record(); torch.cuda.synchronize()
    print(f"config-1 fused fp32 train step: {a.elapsed_time(b) / 50:.4f} ms")
print("done")
