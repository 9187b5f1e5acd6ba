"""Probe: time the one-launch scatter-add at the bench shape (CUDA events).  RB_SCATTER_GRID overrides the grid."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from recboard_b200 import ops, synth  # noqa: E402

N, D, B, S = 1_000_000, 128, 4096, 50
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
seqs = synth.sequences(B, S, N, g, dev)
go = synth.embeddings(B * S, D, g, dev, torch.bfloat16, gain=0.01)
for dt in (torch.float32, torch.bfloat16):
    table = torch.zeros(N + 1, D, dtype=dt, device=dev)
    for _ in range(3):
        ops.scatter_add_rows_(table, go, seqs.view(-1), padding_idx=0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ops.scatter_add_rows_(table, go, seqs.view(-1), padding_idx=0)
    b.record(); torch.cuda.synchronize()
    print(dt, "scatter_add ms", a.elapsed_time(b) / 20)
