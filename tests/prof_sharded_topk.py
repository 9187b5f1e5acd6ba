"""Probe (torchrun, N ranks): where the time of sharded_topk goes on a 10M-item table (d=256, K=100)."""
import os, sys, ctypes
from pathlib import Path
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from recboard_b200 import ops, sharded, synth, _lib as L  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
n_total, rows, K = 10_000_000, 4096, 100
a, b = sharded.shard_bounds(n_total, world, rank)
g = torch.Generator(device=dev).manual_seed(77 + rank)
gq = torch.Generator(device=dev).manual_seed(78)
Wn = ops.normalize_rows(synth.embeddings(b - a, 256, g, dev, torch.bfloat16), out_dtype=torch.bfloat16)
Un = ops.normalize_rows(synth.embeddings(rows, 256, gq, dev, torch.float32), out_dtype=torch.bfloat16)
crow, col = synth.seen_csr(rows, n_total, gq, dev)


def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_local = t(lambda: ops.topk_eval(Un, Wn, K, crow, col, id_base=a))
vals, ids = ops.topk_eval(Un, Wn, K, crow, col, id_base=a)
t_gather = t(lambda: sharded.allgather_topk(vals, ids))
av, ai = sharded.allgather_topk(vals, ids)
t_merge = t(lambda: ops.topk_merge(av, ai))
t_all = t(lambda: sharded.sharded_topk(Un, Wn, K, a, crow, col))
out = (ctypes.c_int64 * 8)()
L.check(L.lib().rb_topk_debug_layout(rows, b - a, 256, K, L.MODE_BF16, col.numel(), out), "layout")
n_sub, cap, n0, off_lad, off_cnt, off_ovf, off_cand, used = list(out)
ws = max(L.Workspace._bufs.values(), key=lambda x: x.numel())
cnt = ws[off_cnt:off_cnt + 4 * rows * n_sub].view(torch.int32).view(rows, n_sub)
ovf = ws[off_ovf:off_ovf + 4 * rows].view(torch.int32)
print(f"[rank {rank}] local topk {t_local:.3f} ms, all-gather {t_gather:.3f}, merge {t_merge:.3f}, sharded_topk {t_all:.3f}; "
      f"n_sub {n_sub} cap {cap} cand/row mean {cnt.sum(1).float().mean():.0f} max sub-list {int(cnt.max())} overflow rows {int(ovf.sum())}", flush=True)
dist.destroy_process_group()
