"""Probe: time the top-K pipeline (CUDA events).  usage: prof_topk.py [N] [d] [B] [K] [reps]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from recboard_b200 import ops, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
M = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
K = int(sys.argv[4]) if len(sys.argv) > 4 else 50
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
U = synth.embeddings(M, D, g, dev, torch.bfloat16, gain=1.5)
W = synth.embeddings(N, D, g, dev, torch.bfloat16, gain=1.5)
crow, col = synth.seen_csr(M, N, g, dev)
for _ in range(3):
    ops.topk_eval(U, W, K, crow, col)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    ops.topk_eval(U, W, K, crow, col)
b.record(); torch.cuda.synchronize()
print("topk_eval ms", a.elapsed_time(b) / reps)

# ---- what the candidate sweep left behind (rb_topk_debug_layout): candidates per row, overflowed rows, ladder state
import ctypes
from recboard_b200 import _lib as L
out = (ctypes.c_int64 * 8)()
L.check(L.lib().rb_topk_debug_layout(M, N, D, K, L.MODE_BF16, col.numel(), out), "rb_topk_debug_layout")
n_sub, cap, n0_tiles, off_lad, off_cnt, off_ovf, off_cand, used = list(out)
ws = next(iter(L.Workspace._bufs.values()))
cnt = ws[off_cnt:off_cnt + 4 * M * n_sub].view(torch.int32).view(M, n_sub)
ovf = ws[off_ovf:off_ovf + 4 * M].view(torch.int32)
lad = ws[off_lad:off_lad + 64 * M].view(torch.int32).view(M, 16)
per_row = cnt.sum(1).float()
print(f"n_sub {n_sub} cand_cap {cap} prefix tiles {n0_tiles}; candidates/row mean {per_row.mean():.1f} max {per_row.max():.0f}; "
      f"max sub-list {int(cnt.max())}; overflow rows {int(ovf.sum())}; ladder counters (mean) {lad[:, 8:].float().mean(0).tolist()}")
