"""Profiling target: config-4 retrieval (d = 256, cosine, top-100) on a table shard, B = 256 and B = 4096.
Usage: python tests/prof_config4.py [n_items] [--time]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from recboard_b200 import ops, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1_250_000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(4)
W = ops.normalize_rows(synth.embeddings(N, 256, g, dev, torch.float32), out_dtype=torch.bfloat16)
U = ops.normalize_rows(synth.embeddings(4096, 256, g, dev, torch.float32), out_dtype=torch.bfloat16)
crow, col = synth.seen_csr(4096, N, g, dev)
for B in (256, 4096):
    c = crow[:B + 1].contiguous(); cc = col[:int(c[-1])].contiguous(); Ub = U[:B].contiguous()
    for _ in range(3):
        ops.topk_eval(Ub, W, 100, c, cc)
    torch.cuda.synchronize()
    if "--time" in sys.argv:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            ops.topk_eval(Ub, W, 100, c, cc)
        b.record(); torch.cuda.synchronize()
        print(f"B={B} N={N} d=256 top-100: {a.elapsed_time(b) / 20:.4f} ms  ({N * 512 / (a.elapsed_time(b) / 20 * 1e-3) / 1e12:.2f} TB/s of table reads)")
print("done")
